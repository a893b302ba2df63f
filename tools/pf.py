import sys; sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
for name in ('cfg4', 'cfg2', 'cfg3'):
    for variant in (1, 3):
        sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
        sim.set_option("step_kernel_variant", variant)
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
            print(f"{name} variant {variant}: {K}-tick {ms*1e3:.1f} us -> {n/ms*1e3/1e9:.2f} G steps/s", flush=True)
        sim.close()
