"""Fused actor-critic forward (hlynr_policy_forward) vs the torch module (fp32 / TF32 / bf16 autocast) at rollout sizes."""
import sys, os; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200.policy import FusedActorCritic, ReferenceActorCritic
net = ReferenceActorCritic(device="cuda")
fused = FusedActorCritic(net)
FLOP_PER_ROW = 2 * (104 * 512 + 512 * 512 + 512 * 256 + 256 * 7)
def timed(f, reps):
    f(); torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e3
for n in (4096, 16384, 131072, 1 << 20):
    obs = torch.randn(n, 104, device="cuda")
    line = f"n={n}:"
    for cl in (1, 2, 4):
        fused.set_option("cluster", cl)
        us = timed(lambda: fused(obs), 20)
        line += f" fused cluster {cl}: {us:.1f} us ({n * FLOP_PER_ROW / us / 1e6:.0f} TFLOP/s);"
    with torch.no_grad():
        torch.backends.cuda.matmul.allow_tf32 = False
        line += f" torch fp32 {timed(lambda: net(obs), 5):.1f} us"
        torch.backends.cuda.matmul.allow_tf32 = True
        line += f"; torch tf32 {timed(lambda: net(obs), 10):.1f} us"
        with torch.autocast("cuda", dtype=torch.bfloat16):
            line += f"; torch bf16 autocast {timed(lambda: net(obs), 10):.1f} us"
    print(line, flush=True)
