import sys; sys.path.insert(0, '.'); sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__)))); sys.path.insert(0, 'tests')
import numpy as np
from common import CudaBatch, golden_setup, load_golden
g = load_golden("volley3_cfg4_f32_zem")
P, cur = golden_setup(g)
meta = g["meta"]
cuda = CudaBatch(P, cur, meta["n_envs"], seed=meta["seed"], float64=False)
obs0 = cuda.reset()
e = 3
np.set_printoptions(linewidth=250, precision=5, suppress=True)
first = None
for t in range(1340):
    obs, rew, te, tr, tobs, info = cuda.step(g["actions"][t])
    d = np.abs(obs[e] - g["obs"][t][e])
    fl = (info["flags"][e], g["flags"][t][e])
    if (d.max() > 2e-5 and t > 1000 and (first is None or t < first + 12)) or t in (1326, 1327):
        if first is None: first = t
        print(t, "maxdiff", d.max(), "ch", d.argmax(), "flags", fl, "dist", info["distance"][e], g["distance"][t][e], "mi", info["missiles_intercepted"][e], g["missiles_intercepted"][t][e])
        print("   cuda", obs[e][[0,1,2,3,4,5,13,14,15,16,24,25]])
        print("   gold", g["obs"][t][e][[0,1,2,3,4,5,13,14,15,16,24,25]])
