import sys, os; sys.path.insert(0, "/root/repo")
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
sim = HlynrSim(config.baseline_config("cfg4"), n_envs=n, warn_dead=False)
sim.reset()
prev = 0.0
for w in range(40):
    sim.rollout(250, None, want_obs=False)
    ep = sim.stats()["episodes"]
    print(f"ticks {250*(w+1):5d}: finished episodes per tick {(ep - prev) / 250:.0f}", flush=True)
    prev = ep
