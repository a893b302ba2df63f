"""Host-path step time (numpy VecEnv API) vs chunk schedule, observation width and action buffer kind at 2^20 envs."""
import sys, os, time; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1 << 20
cfg = config.baseline_config("cfg4")
rng = np.random.default_rng(0)
for od in (26, 17):
    v = HlynrVecEnv(cfg, n_envs=n, seed=1, warn_dead=False, lazy_infos=True, obs_dim=od)
    v.reset(); v.sim.rollout(1200, None, want_obs=False)
    pageable = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
    pinned = [torch.from_numpy(a).pin_memory().numpy() for a in pageable]
    for label, chunks, growth in (("uniform16", 16, 0), ("uniform8", 8, 0), ("uniform24", 24, 0), ("geo x2", 0, 16), ("geo x1.5", 0, 12), ("geo x3", 0, 24)):
        v.sim.set_option("host_chunks", chunks); v.sim.set_option("host_chunk_growth", growth)
        out = []
        for acts in (pageable, pinned):
            for k in range(3): v.step(acts[k % 2])
            t0 = time.perf_counter()
            for k in range(20): v.step(acts[k % 2])
            out.append((time.perf_counter() - t0) / 20 * 1e3)
        print(f"obs_dim {od} {label:10s}: pageable actions {out[0]:.3f} ms, pinned actions {out[1]:.3f} ms per step", flush=True)
    v.close()
