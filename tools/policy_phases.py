"""Per-phase timing of the fused policy kernel (option "timing": SM-clock timestamps of CTA 0's tiles)."""
import sys, os, ctypes as C; sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from hlynr_intercept_b200 import _lib
from hlynr_intercept_b200.policy import FusedActorCritic, ReferenceActorCritic
n = int(sys.argv[1]) if len(sys.argv) > 1 else 131072
fused = FusedActorCritic(ReferenceActorCritic(device="cuda"))
fused.set_option("timing", 1)
obs = torch.randn(n, 104, device="cuda")
names = ["x load", "L1 mma", "L1 epi", "L2 mma", "L2 epi", "L3 mma", "L3 epi", "head mma", "head epi"]
for cl in (1, 2):
    fused.set_option("cluster", cl)
    for _ in range(3): fused(obs)
    t = np.zeros((8, 16), np.int64)
    _lib.check(fused.L.hlynr_policy_get_timing(fused.h, t.ctypes.data_as(C.c_void_p)))
    tiles = [r for r in t if r[9] > 0]
    d = np.array([[r[k + 1] - r[k] for k in range(9)] for r in tiles], float) / 1.9e3   # us at ~1.9 GHz
    print(f"cluster {cl}: {len(tiles)} tiles of CTA 0; per-tile us (mean over tiles, tile 0 excluded): " +
          ", ".join(f"{nm} {x:.2f}" for nm, x in zip(names, d[1:].mean(axis=0))) + f"; total {d[1:].sum(axis=1).mean():.1f} us; "
          f"tile-to-tile gap {np.mean([(tiles[k + 1][0] - tiles[k][9]) / 1.9e3 for k in range(len(tiles) - 1)]):.2f} us")
