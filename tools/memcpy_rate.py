"""Host memcpy rates on the box: pageable -> pinned, 1..8 threads (numpy copyto releases the GIL)."""
import numpy as np, torch, time, threading
n = 1 << 20
src = np.random.rand(n, 6).astype(np.float32)
pin = torch.empty(n, 6).pin_memory().numpy()
page = np.empty_like(src)
def run(dst, T, K=20):
    per = n // T
    def work(k):
        np.copyto(dst[k * per:(k + 1) * per], src[k * per:(k + 1) * per])
    best = 1e9
    for _ in range(K):
        th = [threading.Thread(target=work, args=(k,)) for k in range(T)]
        t0 = time.perf_counter()
        for t in th: t.start()
        for t in th: t.join()
        best = min(best, time.perf_counter() - t0)
    return best
for T in (1, 2, 4, 8):
    a, b = run(pin, T), run(page, T)
    print(f"{T} threads: pageable->pinned {a*1e3:.2f} ms ({src.nbytes/a/1e9:.1f} GB/s) | pageable->pageable {b*1e3:.2f} ms ({src.nbytes/b/1e9:.1f} GB/s)")
