"""Raw pinned-memory PCIe rates of the box (the floor of the host path): D2H 116 MB, H2D 25 MB, and both at once."""
import torch, time
n = 1 << 20
d_obs = torch.empty(n * 110, dtype=torch.uint8, device='cuda'); h_obs = torch.empty(n * 110, dtype=torch.uint8).pin_memory()
d_act = torch.empty(n * 24, dtype=torch.uint8, device='cuda'); h_act = torch.empty(n * 24, dtype=torch.uint8).pin_memory()
s1, s2 = torch.cuda.Stream(), torch.cuda.Stream()
def t(f, K=20):
    f(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(K): f()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / K
a = t(lambda: h_obs.copy_(d_obs, non_blocking=True)); print(f"D2H 115 MB: {a*1e3:.2f} ms = {n*110/a/1e9:.1f} GB/s")
b = t(lambda: d_act.copy_(h_act, non_blocking=True)); print(f"H2D 25 MB: {b*1e3:.2f} ms = {n*24/b/1e9:.1f} GB/s")
def both():
    with torch.cuda.stream(s1): h_obs.copy_(d_obs, non_blocking=True)
    with torch.cuda.stream(s2): d_act.copy_(h_act, non_blocking=True)
c = t(both); print(f"both directions: {c*1e3:.2f} ms")
def chunked():
    per = n * 110 // 8
    for k in range(8):
        with torch.cuda.stream(s1 if k % 2 else s2): h_obs[k*per:(k+1)*per].copy_(d_obs[k*per:(k+1)*per], non_blocking=True)
d = t(chunked); print(f"D2H in 8 chunks over 2 streams: {d*1e3:.2f} ms")
