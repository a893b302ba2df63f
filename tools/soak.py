"""Long fused-rollout soak at 2^20 envs: finite observations inside the Box bounds, and exact bookkeeping identities
(env_steps = n x ticks; sum of finished-episode lengths + steps of the running episodes = env_steps)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
for name in ("cfg4", "cfg3", "cfg2"):
    sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
    sim.reset()
    ticks = 0
    for rep in range(6):
        obs, rsum, dcount = sim.rollout(5000, None)
        ticks += 5000
        torch.cuda.synchronize()
        assert torch.isfinite(obs).all() and obs.min().item() >= -2.0 and obs.max().item() <= 1.0, (name, rep)
    st = sim.stats()
    running = int(torch.as_tensor(sim.export_state(0, n)["steps"]).sum())
    assert st["env_steps"] == float(n) * ticks, (st["env_steps"], n * ticks)
    assert st["length_sum"] + running == st["env_steps"], (st["length_sum"], running, st["env_steps"])
    causes = sum(st[k] for k in ("hit_target", "interceptor_crash", "fuel_out", "missile_ground", "worsening", "timeouts")) + st["successes"]
    print(f"{name}: {ticks} ticks x {n} envs ok; episodes {st['episodes']:.0f}, successes {st['successes']:.0f}, "
          f"termination causes + successes {causes:.0f}, mean length {st['length_sum'] / st['episodes']:.1f}", flush=True)
    sim.close()
