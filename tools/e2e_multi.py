"""Where does the host path's time go with one rank per GPU?  torchrun --nproc-per-node N tools/e2e_multi.py
Phases: each rank alone / all ranks together / pinned caller actions / after binding the rank to its GPU's NUMA node."""
import sys, os, time, subprocess; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch, torch.distributed as dist
from hlynr_intercept_b200 import config
from hlynr_intercept_b200.vec_env import HlynrVecEnv
rank, world, lr = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
dist.init_process_group("gloo")
torch.cuda.set_device(lr)
n = 1 << 20
if rank == 0:
    for cmd in (["nvidia-smi", "topo", "-m"], ["lscpu"]):
        out = subprocess.run(cmd, capture_output=True, text=True).stdout
        print("\n".join(l for l in out.split("\n") if cmd[0] != "lscpu" or "NUMA" in l or "CPU(s)" in l or "Model name" in l), flush=True)
def make():
    v = HlynrVecEnv(config.baseline_config("cfg4"), n_envs=n, device=lr, seed=99, env_id_offset=rank * n, warn_dead=False, lazy_infos=True)
    v.reset(); v.sim.rollout(1200, None, want_obs=False)
    return v
def run(v, acts, K):
    for k in range(2): v.step(acts[k % 2])
    t0 = time.perf_counter()
    for k in range(K): v.step(acts[k % 2])
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / K * 1e3
def gather(x):
    t = torch.tensor([x], dtype=torch.float64); l = [torch.zeros(1, dtype=torch.float64) for _ in range(world)]
    dist.all_gather(l, t); return [round(float(a), 2) for a in l]
def phase(name, v, acts, K=12, solo=False):
    if solo:
        res = 0.0
        for r in range(world):
            dist.barrier()
            if r == rank: res = run(v, acts, K)
        dist.barrier()
    else:
        dist.barrier(); res = run(v, acts, K); dist.barrier()
    g = gather(res)
    if rank == 0: print(f"{name}: ms/step per rank {g}  -> aggregate {sum(n / (m * 1e-3) for m in g) / 1e6:.0f} M env-steps/s", flush=True)
rng = np.random.default_rng(rank)
acts = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(2)]
v = make()
print(f"rank {rank}: affinity {sorted(os.sched_getaffinity(0))[:4]}.. ({len(os.sched_getaffinity(0))} cpus)", flush=True)
phase("solo, unpinned actions", v, acts, solo=True)
phase("all ranks, unpinned actions", v, acts)
phase("all ranks, pinned actions", v, [v._act, v._act])
for th in (2, 4):
    v.sim.set_option("host_threads", th)
    phase(f"all ranks, unpinned actions, host_threads {th}", v, acts)
v.close()
try:
    import pynvml
    pynvml.nvmlInit()
    h = pynvml.nvmlDeviceGetHandleByIndex(lr)
    words = pynvml.nvmlDeviceGetCpuAffinity(h, (os.cpu_count() + 63) // 64)
    cpus = [64 * i + b for i, w in enumerate(words) for b in range(64) if (w >> b) & 1]
    os.sched_setaffinity(0, cpus)
    print(f"rank {rank}: bound to {len(cpus)} cpus {cpus[:4]}..", flush=True)
except Exception as e:
    print(f"rank {rank}: no NUMA binding ({e})", flush=True)
v = make()
phase("NUMA-bound, all ranks, unpinned actions", v, acts)
phase("NUMA-bound, all ranks, pinned actions", v, [v._act, v._act])
v.close()
dist.destroy_process_group()
