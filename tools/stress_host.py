"""Stress of the staging thread pool: 1500 steps through hlynr_step_host with pageable actions (8 staging threads, 32 chunks),
every step compared bit for bit with the device API on a twin simulator."""
import sys, os, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = 200001
ref = HlynrSim(config.baseline_config("cfg4"), n_envs=n, seed=77, warn_dead=False)
v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, seed=77, warn_dead=False, lazy_infos=True, copy_outputs=False)
v.sim.set_option("host_threads", 8); v.sim.set_option("host_chunks", 32)
ref.reset(); v.reset(); ref.rollout(900, None); v.sim.rollout(900, None)
rng = np.random.default_rng(1)
pool = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(7)]
dev = [torch.as_tensor(a).cuda() for a in pool]
for t in range(1500):
    k = t % 7
    if t % 97 == 0: v.sim.set_option("host_chunks", int(rng.integers(1, 40)))
    o, r, te, tr, _, _ = ref.step(dev[k], want_terminal_obs=False)
    obs, rew, dones, infos = v.step(pool[k])
    assert (obs == o.cpu().numpy()).all() and (rew == r.cpu().numpy()).all() and (dones == (te | tr).cpu().numpy().astype(bool)).all(), t
print("stress ok: 1500 steps x", n, "envs bit-identical")
