import sys, time; sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch, numpy as np
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
for name in ['cfg4','cfg2','cfg3']:
    for prec in ['fp32','fp64']:
        n = 1<<20
        sim = HlynrSim(config.baseline_config(name), n_envs=n, precision=prec, warn_dead=False)
        sim.reset()
        act = (torch.rand(n,6,device='cuda')*2-1)
        for _ in range(5): sim.step(act)
        torch.cuda.synchronize()
        e0=torch.cuda.Event(enable_timing=True); e1=torch.cuda.Event(enable_timing=True)
        K=50
        e0.record()
        for _ in range(K): sim.step(act, want_terminal_obs=False)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1)/K
        e0.record(); sim.rollout(64, None, want_obs=False); e1.record(); torch.cuda.synchronize()
        msr = e0.elapsed_time(e1)/64
        print(f"{name} {prec}: step {ms*1e3:.1f} us -> {n/ms*1e3/1e9:.2f} G steps/s ; fused64 {msr*1e3:.1f} us/step -> {n/msr*1e3/1e9:.2f} G steps/s", flush=True)
        sim.close()
