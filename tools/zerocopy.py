"""Does the step kernel writing its outputs STRAIGHT into pinned host memory (UVA, posted PCIe writes from the SMs) beat
kernel -> HBM -> copy-engine D2H?  Times hlynr_step with host-resident output pointers, whole shard and per chunk size."""
import sys, os, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from hlynr_intercept_b200 import config, _lib
from hlynr_intercept_b200.sim import HlynrSim
n = 1 << 20
sim = HlynrSim(config.baseline_config("cfg4"), n_envs=n, warn_dead=False)
sim.reset(); sim.rollout(1200, None, want_obs=False)
act = torch.rand(n, 6, device='cuda') * 2 - 1
h_obs = torch.empty(n, 26).pin_memory(); h_rew = torch.empty(n).pin_memory()
h_te = torch.empty(n, dtype=torch.uint8).pin_memory(); h_tr = torch.empty(n, dtype=torch.uint8).pin_memory()
d = sim._alloc_out()
p = lambda t: C.c_void_p(t.data_ptr())
def step(obs, rew, te, tr):
    _lib.check(sim.L.hlynr_step(sim.h, p(act), p(obs), p(rew), p(te), p(tr), None, None, 1, sim._stream()))
def timed(f, K=30):
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K): f()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / K
a = timed(lambda: step(d["obs"], d["reward"], d["terminated"], d["truncated"]))
print(f"all outputs in HBM: {a*1e3:.1f} us")
b = timed(lambda: step(d["obs"], h_rew, h_te, h_tr))
print(f"reward/terminated/truncated written to pinned host memory: {b*1e3:.1f} us")
c = timed(lambda: step(h_obs, h_rew, h_te, h_tr))
print(f"all outputs written to pinned host memory: {c*1e3:.1f} us = {n*110/c/1e6:.1f} GB/s over PCIe")
def copy_path():
    step(d["obs"], d["reward"], d["terminated"], d["truncated"])
    h_obs.copy_(d["obs"], non_blocking=True); h_rew.copy_(d["reward"], non_blocking=True)
    h_te.copy_(d["terminated"], non_blocking=True); h_tr.copy_(d["truncated"], non_blocking=True)
e = timed(copy_path)
print(f"kernel + 4 copy-engine D2H copies, one stream, no chunking: {e*1e3:.1f} us")
ok = bool((h_obs == d["obs"].cpu()).all())
step(h_obs, h_rew, h_te, h_tr); torch.cuda.synchronize()
step(d["obs"], d["reward"], d["terminated"], d["truncated"]); torch.cuda.synchronize()
print("copy path consistent:", ok)
sim.close()
