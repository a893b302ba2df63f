"""Small run that touches every kernel path (direct / TMA / pipe step kernels, fused rollout, reset, host path, post pipeline,
GAE) at ragged sizes.  Written for `compute-sanitizer --tool memcheck|racecheck python tools/sanitize.py`; compute-sanitizer is
closed on this GPU pool, so here it only serves as an all-paths smoke run (bounds are covered by the ragged-size parity tests)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
from hlynr_intercept_b200.vec_env import HlynrVecEnv
import sweep_configs
n = 1000   # ragged: not a multiple of 32 or 128
for name, prec in (("cfg4", "fp32"), ("cfg2", "fp32"), ("cfg3", "fp32"), ("cfg3", "fp64"), ("cfg4", "fp64")):
    sim = HlynrSim(config.baseline_config(name), n_envs=n, precision=prec, warn_dead=False)
    sim.reset()
    sim.rollout(40, None)
    for k in range(6):
        sim.step(torch.rand(n, 6, device="cuda") * 2 - 1, want_info=True)
    sim.reset(torch.randint(0, 2, (n,), dtype=torch.uint8))
    torch.cuda.synchronize(); sim.close()
for k in (1, 4, 12, 27):   # volley / los / body / spherical / DR mixes
    sim = HlynrSim(sweep_configs.sweep_config(k), n_envs=n, warn_dead=False)
    sim.reset(); sim.rollout(30, None)
    for _ in range(4): sim.step(torch.rand(n, 6, device="cuda") * 2 - 1, want_info=True)
    torch.cuda.synchronize(); sim.close()
v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=40001, warn_dead=False, lazy_infos=True)
v.sim.set_option("host_chunks", 5)
v.reset(); v.sim.rollout(900, None)
for _ in range(4): v.step(np.random.uniform(-1, 1, (40001, 6)).astype(np.float32))
v.close()
from hlynr_intercept_b200.post import HlynrObsPipeline
from hlynr_intercept_b200.rollout import DeviceRolloutCollector, GaussianMlpPolicy
sim = HlynrSim(config.baseline_config("cfg4"), n_envs=n, warn_dead=False)
sim.reset(); sim.rollout(950, None)   # close to the first terminations: the collector sees finished episodes and time-outs
pipe = HlynrObsPipeline(sim, n_stack=4)
col = DeviceRolloutCollector(pipe, GaussianMlpPolicy(104, device=sim.device), 6)
col.collect(); col.collect()
torch.cuda.synchronize()
pipe.close(); sim.close()
print("sanitize run complete")
