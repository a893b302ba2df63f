"""Throughput of step + on-device frame-stack/normalise at 2^20 envs (scratch script for gpurun)."""
import sys
sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
from hlynr_intercept_b200.post import HlynrObsPipeline
n = 1 << 20
for k in (4, 1):
    for training in (True, False):
        sim = HlynrSim(config.baseline_config("cfg4"), n_envs=n, warn_dead=False)
        pipe = HlynrObsPipeline(sim, n_stack=k, training=training, want_terminal_obs=True)
        pipe.reset()
        act = torch.rand(n, 6, device='cuda') * 2 - 1
        for _ in range(10): pipe.step(act)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        K = 200
        e0.record()
        for _ in range(K): pipe.step(act)
        e1.record(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / K
        print(f"n_stack {k} training {training}: step+post {ms*1e3:.1f} us -> {n/ms*1e3/1e9:.2f} G env-steps/s", flush=True)
        pipe.close(); sim.close()
