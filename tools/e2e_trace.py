import sys, os, time, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config, _lib
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = 1 << 20
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, seed=1, warn_dead=False, lazy_infos=True)
v.reset(); v.sim.rollout(1200, None, want_obs=False)
for k in range(6): v.step(acts[k % 2])
print("--- pinned", file=sys.stderr, flush=True)
for k in range(4): v.step(v._act)
v.close()
