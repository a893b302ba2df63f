import sys; sys.path.insert(0, '.')
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
for rep in range(2):
    for pw in (0, 1, 2, 3):
        sim = HlynrSim(config.baseline_config('cfg4'), n_envs=n, warn_dead=False)
        sim.set_option("prefetch_waves", pw)
        sim.reset()
        pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
        for k in range(200): sim.step(pool[k % 4], want_terminal_obs=False)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = 2000
        e0.record()
        for k in range(K): sim.step(pool[k % 4], want_terminal_obs=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"cfg4 prefetch_waves {pw}: 2000-tick bench {ms*1e3:.1f} us -> {n/ms*1e3/1e9:.2f} G steps/s", flush=True)
        sim.close()
