"""API-mode step time at the sizes BASELINE.json names: cfg2 at 4096 envs, cfg3 at 262144 envs, cfg4 at 2^20 (eager launches and,
for the launch-bound small batch, a CUDA graph of 100 ticks)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
def timed(f, K):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K): f(k)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
for name, n in (("cfg2", 4096), ("cfg3", 262144), ("cfg4", 1 << 20)):
    sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
    sim.reset(); sim.rollout(1200, None, want_obs=False)
    pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
    step = lambda k: sim.step(pool[k % 4], want_terminal_obs=False)
    timed(step, 50)
    us = timed(step, 1000)
    line = f"{name} n={n}: {us:.2f} us per step (eager) -> {n / us / 1e3:.3f} G env-steps/s"
    if n <= 65536:
        s = torch.cuda.Stream()
        with torch.cuda.stream(s):
            for k in range(3): step(k)
            g = torch.cuda.CUDAGraph()
            period = sim.ring_period() if hasattr(sim, "ring_period") else None
            try:
                with torch.cuda.graph(g, stream=s):
                    for k in range(100): step(k)
                torch.cuda.synchronize()
                e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
                e0.record(s)
                for _ in range(20): g.replay()
                e1.record(s); torch.cuda.synchronize()
                gus = e0.elapsed_time(e1) / 2000 * 1e3
                line += f"; CUDA graph of 100 ticks (timing only: the ring row is baked into a graph, see hlynr_ring_period): {gus:.2f} us per step -> {n / gus / 1e3:.3f} G env-steps/s"
            except Exception as ex:
                line += f"; graph capture failed: {ex}"
    print(line, flush=True)
    sim.close()
