"""API-mode step time early in the episode (all envs in phase) and at the steady state (episodes desynchronised by a fused
2000-tick rollout, ~1000 auto-resets per tick at 2^20 envs).  HLYNR_B200_LIB selects the build of the library (A/B tests).
  python tools/aged_time.py cfg4,cfg2,cfg3 [fp32|fp64] [n_envs]"""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
names = sys.argv[1].split(',') if len(sys.argv) > 1 else ['cfg4']
precision = sys.argv[2] if len(sys.argv) > 2 else 'fp32'
n = int(sys.argv[3]) if len(sys.argv) > 3 else 1 << 20
tag = os.path.basename(os.environ.get("HLYNR_B200_LIB", "default")) + " " + os.environ.get("HLYNR_OPTS", "")
def timed(sim, pool, K):
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for k in range(K): sim.step(pool[k % 4], want_terminal_obs=False)
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K * 1e3
for name in names:
    sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False, precision=precision)
    for kv in filter(None, os.environ.get("HLYNR_OPTS", "").split(",")):   # e.g. HLYNR_OPTS=pdl=1
        sim.set_option(kv.split("=")[0], int(kv.split("=")[1]))
    sim.reset()
    pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
    timed(sim, pool, 50)
    early = timed(sim, pool, 150)
    sim.rollout(2000, None, want_obs=False)
    timed(sim, pool, 20)
    aged = min(timed(sim, pool, 300) for _ in range(3))
    print(f"{tag} {name} {precision} n={n}: early {early:.1f} us, steady state {aged:.1f} us -> {n / aged / 1e3:.2f} G steps/s", flush=True)
    sim.close()
