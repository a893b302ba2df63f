"""API-mode step time early in the episode (all envs in phase) and at the steady state (episodes desynchronised by a fused
1500-tick rollout, ~900 auto-resets per tick at 2^20 envs), per step-kernel variant.  HLYNR_B200_LIB selects the build."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['cfg4']
variants = [int(v) for v in sys.argv[2].split(',')] if len(sys.argv) > 2 else [1]
tag = os.path.basename(os.environ.get("HLYNR_B200_LIB", "default"))
def timed(sim, pool, K):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K): sim.step(pool[k % 4], want_terminal_obs=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
for name in names:
    for variant in variants:
        sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False)
        sim.set_option("step_kernel_variant", variant)
        sim.reset()
        pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
        timed(sim, pool, 50)
        early = timed(sim, pool, 150)
        sim.rollout(1500, None, want_obs=False)
        timed(sim, pool, 20)
        aged = timed(sim, pool, 300)
        print(f"{tag} {name} variant {variant}: early {early:.1f} us, steady state {aged:.1f} us -> {n / aged / 1e3:.2f} G steps/s", flush=True)
        sim.close()
