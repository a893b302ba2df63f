"""A few API-mode ticks at the steady state of the bench (episodes desynchronised by a 1500-tick fused rollout first):
the launch an ncu capture should look at (`ncu -k regex:step_kernel -s 6 -c 1`: launch 0 is the fused rollout)."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
name = sys.argv[1] if len(sys.argv) > 1 else 'cfg4'
sim = HlynrSim(config.baseline_config(name), n_envs=n, warn_dead=False, precision=os.environ.get('HLYNR_PRECISION', 'fp32'))
for kv in filter(None, os.environ.get('HLYNR_OPTS', '').split(',')):   # e.g. HLYNR_OPTS=split=1
    sim.set_option(kv.split('=')[0], int(kv.split('=')[1]))
sim.reset()
sim.rollout(1500, None, want_obs=False)
pool = [(torch.rand(n, 6, device='cuda') * 2 - 1) for _ in range(4)]
for k in range(12): sim.step(pool[k % 4], want_terminal_obs=False)
torch.cuda.synchronize()
print(sim.stats())
sim.close()
