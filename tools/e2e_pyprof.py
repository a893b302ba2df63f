import sys, os, time, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config, _lib
from hlynr_intercept_b200.vec_env import HlynrVecEnv
n = 1 << 20
rng = np.random.default_rng(0)
acts = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, seed=1, warn_dead=False, lazy_infos=True)
v.reset(); v.sim.rollout(1200, None, want_obs=False)
for k in range(4): v.step(acts[k % 2])
T = dict(async_=0, ccall=0, lor=0, infos=0); K = 30
p = lambda x: x.ctypes.data_as(C.c_void_p)
for k in range(K):
    t0 = time.perf_counter(); v.step_async(acts[k % 2]); t1 = time.perf_counter()
    a, v._pending = v._pending, None
    _lib.check(v.sim.L.hlynr_step_host(v.sim.h, p(a), p(v._obs), p(v._rew), p(v._term), p(v._trunc), None, 1)); t2 = time.perf_counter()
    d = np.logical_or(v._term, v._trunc); t3 = time.perf_counter()
    inf = v._build_infos(); t4 = time.perf_counter()
    T["async_"] += t1 - t0; T["ccall"] += t2 - t1; T["lor"] += t3 - t2; T["infos"] += t4 - t3
print({k: round(x / K * 1e3, 3) for k, x in T.items()})
import cProfile, pstats
pr = cProfile.Profile(); pr.enable()
for k in range(20): v.step(acts[k % 2])
pr.disable(); pstats.Stats(pr).sort_stats("cumulative").print_stats(14)
v.close()
