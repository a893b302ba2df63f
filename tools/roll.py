"""Rollout collection eager vs CUDA graph at several batch sizes (scratch script for gpurun)."""
import sys; sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
from hlynr_intercept_b200.post import HlynrObsPipeline
from hlynr_intercept_b200.rollout import DeviceRolloutCollector, GaussianMlpPolicy
torch.backends.cuda.matmul.allow_tf32 = True
for n in (4096, 16384, 131072):
    for arch in ((64, 64), (512, 512, 256)):
        sim = HlynrSim(config.baseline_config("cfg4"), n_envs=n, warn_dead=False)
        pipe = HlynrObsPipeline(sim, n_stack=4, training=True)
        pol = GaussianMlpPolicy(104, net_arch=arch)
        col = DeviceRolloutCollector(pipe, pol, n_steps=48)
        col.collect(); torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(); col.collect(); col.collect(); e1.record(); torch.cuda.synchronize()
        eager = e0.elapsed_time(e1) / 2
        col.capture()
        col.replay(); torch.cuda.synchronize()
        e0.record(); col.replay(); col.replay(); e1.record(); torch.cuda.synchronize()
        graph = e0.elapsed_time(e1) / 2
        print(f"n={n} arch={arch}: eager {eager/48*1e3:.0f} us/step -> {n*48/eager*1e3/1e6:.1f} M env-steps/s ; graph {graph/48*1e3:.0f} us/step -> {n*48/graph*1e3/1e6:.1f} M env-steps/s", flush=True)
        pipe.close(); sim.close()
