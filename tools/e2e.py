"""e2e sweep of hlynr_step_host: chunks x threads at 2^20 envs (scratch script for gpurun)."""
import sys, time
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = 1 << 20
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
for threads in (1, 4, 8):
    for chunks in (1, 4, 8, 16, 32):
        v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, seed=1, warn_dead=False, lazy_infos=True, copy_outputs=False)
        v.sim.set_option("host_threads", threads)
        v.sim.set_option("host_chunks", chunks)
        v.reset()
        v.sim.rollout(1200, None, want_obs=False)
        for k in range(3):
            v.step(acts[k % 2])
        t0 = time.perf_counter()
        K = 20
        for k in range(K):
            v.step(acts[k % 2])
        dt = (time.perf_counter() - t0) / K
        # pinned actions (no staging copy)
        np.copyto(v._act, acts[0])
        t0 = time.perf_counter()
        for k in range(K):
            v.step(v._act)
        dtp = (time.perf_counter() - t0) / K
        print(f"threads {threads} chunks {chunks}: {dt*1e3:.2f} ms/step -> {n/dt/1e6:.0f} M env-steps/s ; pinned actions {dtp*1e3:.2f} ms", flush=True)
        v.close()
