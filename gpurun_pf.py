import sys; sys.path.insert(0, '.')
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
for name in ('cfg4', 'cfg2'):
    sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
    sim.reset()
    pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
    for k in range(200): sim.step(pool[k % 4], want_terminal_obs=False)
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    for K in (200, 2000):
        e0.record()
        for k in range(K): sim.step(pool[k % 4], want_terminal_obs=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"{name}: {K}-tick {ms*1e3:.1f} us -> {n/ms*1e3/1e9:.2f} G steps/s", flush=True)
    sim.close()
