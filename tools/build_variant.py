"""Builds a variant of the library for A/B timing on the GPU box: python tools/build_variant.py <tag> [-DNAME=VALUE ...]
-> hlynr_intercept_b200/_variants/libhlynr_b200_<tag>.so (git-ignored, travels with gpurun); select it with HLYNR_B200_LIB."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from hlynr_intercept_b200 import build as b
tag, defs = sys.argv[1], sys.argv[2:]
d = os.path.join(b.HERE, "_variants"); os.makedirs(d, exist_ok=True)
so = os.path.join(d, f"libhlynr_b200_{tag}.so")
r = subprocess.run(["nvcc"] + b.NVCC_FLAGS + defs + ["-o", so] + b.SOURCES, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
if r.returncode: sys.exit(r.stdout)
open(so + ".ptxas.txt", "w").write(r.stdout)
print(so)
