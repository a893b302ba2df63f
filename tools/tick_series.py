"""Per-window step time vs finished-episode rate (where does the 200-tick -> 2000-tick slowdown come from?)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
W = 100
sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
sim.reset()
pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
for k in range(50): sim.step(pool[k % 4], want_terminal_obs=False)
torch.cuda.synchronize()
prev = sim.stats()['episodes']
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
for w in range(36):
    e0.record()
    for k in range(W): sim.step(pool[k % 4], want_terminal_obs=False)
    e1.record(); torch.cuda.synchronize()
    ep = sim.stats()['episodes']
    print(f"{name} ticks {50 + w * W:5d}-{50 + (w + 1) * W:5d}: {e0.elapsed_time(e1) / W * 1e3:7.1f} us/launch  done/tick {(ep - prev) / W:8.1f}", flush=True)
    prev = ep
sim.close()
