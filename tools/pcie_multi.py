"""Raw pinned-memory PCIe ceiling of the box with one rank per GPU (torchrun --nproc-per-node N tools/pcie_multi.py):
device->host of the host path's download (111 B/env x 2^20 envs) and host->device of its upload (24 B/env), every rank ALONE and
all ranks TOGETHER.  The aggregate together-rate is the platform's ceiling for the e2e metric, independent of this library."""
import os, sys, time
import torch, torch.distributed as dist
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
torch.cuda.set_device(lr)
n = 1 << 20
d_obs = torch.empty(n * 111, dtype=torch.uint8, device="cuda"); h_obs = torch.empty(n * 111, dtype=torch.uint8).pin_memory()
d_act = torch.empty(n * 24, dtype=torch.uint8, device="cuda"); h_act = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def both():
    with torch.cuda.stream(s1): h_obs.copy_(d_obs, non_blocking=True)
    with torch.cuda.stream(s2): d_act.copy_(h_act, non_blocking=True)
def timed(K=15):
    both(); torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    for _ in range(K): both()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / K
def gather(x):
    l = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(l, torch.tensor([x], dtype=torch.float64)); return [float(a) for a in l]
solo = []
for r in range(world):   # one rank at a time
    dist.barrier()
    if r == rank:
        both(); torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(15): both()
        torch.cuda.synchronize(); mine = (time.perf_counter() - t0) / 15
    dist.barrier()
solo = gather(mine)
together = gather(timed())
if rank == 0:
    gb = n * 111 / 1e9
    print(f"ranks {world}: D2H {n*111/1e6:.0f} MB + H2D {n*24/1e6:.0f} MB per rank per iteration")
    print("alone    ms per iteration:", [round(x * 1e3, 2) for x in solo], "-> D2H GB/s per rank:", [round(gb / x, 1) for x in solo])
    print("together ms per iteration:", [round(x * 1e3, 2) for x in together], "-> D2H GB/s per rank:", [round(gb / x, 1) for x in together])
    print(f"aggregate D2H together: {sum(gb / x for x in together):.1f} GB/s (+ {sum(n*24/1e9 / x for x in together):.1f} GB/s H2D); "
          f"e2e ceiling at {world} ranks: {world * n / max(together) / 1e9:.3f} G env-steps/s ({max(together)*1e3:.2f} ms per step)")
dist.destroy_process_group()
