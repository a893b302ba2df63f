"""Where one HlynrVecEnv.step() spends its time: the C call (hlynr_step_host) vs the Python wrapper around it."""
import sys, os, time, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config, _lib
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = 1 << 20
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, seed=1, warn_dead=False, lazy_infos=True)
v.reset(); v.sim.rollout(1200, None, want_obs=False)
p = lambda x: x.ctypes.data_as(C.c_void_p)
def c_call(a):
    _lib.check(v.sim.L.hlynr_step_host(v.sim.h, p(a), p(v._obs), p(v._rew), p(v._term), p(v._trunc), None, 1))
def t(f, K=30):
    for k in range(3): f(k)
    t0 = time.perf_counter()
    for k in range(K): f(k)
    return (time.perf_counter() - t0) / K * 1e3
for chunks in (8, 16, 24, 32, 48):
    v.sim.set_option("host_chunks", chunks)
    print(f"chunks {chunks}: C call unpinned actions {t(lambda k: c_call(acts[k % 2])):.2f} ms | C call pinned actions {t(lambda k: c_call(v._act)):.2f} ms | "
          f"venv.step unpinned {t(lambda k: v.step(acts[k % 2])):.2f} ms | venv.step pinned {t(lambda k: v.step(v._act)):.2f} ms", flush=True)
v.close()
