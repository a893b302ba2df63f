#!/usr/bin/env python
"""Folds the JSON-lines records the `-m gpu` parity tests append (tests/common.ParityLog -> gpurun_out/parity_report.jsonl)
into profiles/parity_report.json: the worst error per observation channel / output field and the dropped-env count of every
test, plus the worst of each (build, test class).  The tolerances in tests/common.TOL are justified by this file.

  python tools/parity_report.py [gpurun_out/parity_report.jsonl] [profiles/parity_report.json]
"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "parity_report.jsonl")
dst = sys.argv[2] if len(sys.argv) > 2 else os.path.join(ROOT, "profiles", "parity_report.json")

recs = {}
for line in open(src):
    line = line.strip()
    if line:
        r = json.loads(line)
        recs[r["test"]] = r   # the last run of a test wins

LOOSE = [9, 10, 11, 13, 16]
groups = {}
for name, r in sorted(recs.items()):
    cls = name.split("/")[0] + ("/f64" if r.get("f64") else "/f32")
    g = groups.setdefault(cls, {"tests": 0, "dropped": 0, "env_ticks_compared": 0, "obs_abs": [0.0] * 26, "obs_floor": [0.0] * 26,
                                "fields": {}})
    g["tests"] += 1
    g["dropped"] += r["dropped"]
    g["env_ticks_compared"] += r["env_ticks_compared"]
    g["obs_abs"] = [max(a, b) for a, b in zip(g["obs_abs"], r["obs_abs"])]
    g["obs_floor"] = [max(a, b) for a, b in zip(g["obs_floor"], r["obs_floor"])]
    for k, v in r["fields"].items():
        w = g["fields"].setdefault(k, {"abs": 0.0, "floor": 0.0, "rel": 0.0})
        for m in w:
            w[m] = max(w[m], v[m])
for g in groups.values():
    g["worst_obs_well_conditioned"] = max(x for k, x in enumerate(g["obs_abs"]) if k not in LOOSE)
    g["worst_obs_ill_conditioned(9,10,11,13,16)"] = max(g["obs_abs"][k] for k in LOOSE)

out = {"generated_by": "tools/parity_report.py from the records of one `pytest -m gpu` run on B200",
       "columns": {"obs_abs": "worst |cuda - reference| per observation channel (normalised units), over all compared envs and ticks",
                   "obs_floor": "worst (|cuda - reference| - rtol * |reference|): the absolute floor an rtol test would need",
                   "fields.abs/floor/rel": "the same for rewards, info fields and the final state; rel only where |reference| > 1e-3",
                   "dropped": "envs excluded after a decision margin below the build's tolerance AND an actual disagreement"},
       "groups": groups, "tests": list(recs.values())}
os.makedirs(os.path.dirname(dst), exist_ok=True)
json.dump(out, open(dst, "w"), indent=1)
for cls, g in groups.items():
    print(f"{cls:18s} tests {g['tests']:3d} dropped {g['dropped']:3d} obs {g['worst_obs_well_conditioned']:.2e} "
          f"ill {g['worst_obs_ill_conditioned(9,10,11,13,16)']:.2e} reward rel {g['fields'].get('reward', {}).get('rel', 0):.2e} "
          f"floor {g['fields'].get('reward', {}).get('floor', 0):.2e}")
