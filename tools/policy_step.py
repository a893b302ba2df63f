"""A few fused-policy forwards at the rollout batch size: the launches an ncu capture looks at."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200.policy import FusedActorCritic, ReferenceActorCritic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
fused = FusedActorCritic(ReferenceActorCritic(device="cuda"))
if len(sys.argv) > 2: fused.set_option("cluster", int(sys.argv[2]))
obs = torch.randn(n, 104, device="cuda")
for _ in range(4): fused(obs)
torch.cuda.synchronize()
