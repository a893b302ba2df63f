"""TEST INFRASTRUCTURE -- episode-outcome statistics of the UNMODIFIED reference with its OWN random generators
(no draw injection), for the fp32 build's statistical-equivalence test (north_star).  Run in the build container:

    python -m oracle.gen_stat_golden

Writes tests/golden/stat_<cfg>_<policy>.npz: per-episode length, return, minimum distance, final distance and
termination cause of N first episodes under a scripted policy.
"""
import json
import os
import sys

import numpy as np

from hlynr_intercept_b200 import config
from . import ref_harness as rh
from .gen_golden import GOLDEN_DIR

# name: (base cfg or "yaml:<path under rl_system>", policy, episodes, extra env keys, training_step_count)
CASES = {"stat_cfg4_random": ("cfg4", "random", 96, None, None), "stat_cfg4_pursuit": ("cfg4", "pursuit", 96, None, None),
         "stat_cfg2_pursuit": ("cfg2", "pursuit", 96, None, None),
         "stat_hrl_los_pn": ("yaml:configs/hrl/terminal_los.yaml", "los_pn", 96, None, 500000),
         "stat_hrl_rotinv_mixed": ("yaml:configs/hrl/terminal_360_rotinv.yaml", "mixed", 96, None, 1000000),
         "stat_volley3_pursuit": ("cfg4", "pursuit", 64, dict(volley_mode=True, volley_size=3), None)}
CAUSES = ["intercepted", "hit_target", "interceptor_crash", "fuel_out", "missile_ground", "worsening", "timeout"]


def cause_of(info, terminated, env):
    if info["intercepted"]:
        return 0
    if not terminated:
        return 6
    if info["missile_hit_target"]:
        return 1
    if env.interceptor_state["position"][2] < 0:
        return 2
    if env.interceptor_state["fuel"] <= 0:
        return 3
    if env.missile_state["position"][2] <= 0:
        return 4
    return 5


def generate(name):
    rh.install_shim()
    import environment as envmod  # noqa: F401  (the reference module; native numpy RNGs are used here)
    from environment import InterceptEnvironment

    from .gen_golden import reference_yaml_env

    base, policy, n_ep, extra, tsc = CASES[name]
    # the harness may have replaced np in the reference module by the tape proxy: restore real numpy
    envmod.np = np
    cfg = reference_yaml_env(base[5:]) if base.startswith("yaml:") else config.baseline_config(base)
    if extra:
        cfg.update(extra)
    env = InterceptEnvironment(dict(cfg))
    if tsc is not None:
        env.set_training_step_count(tsc)
    pol = {"random": lambda: rh.policy_random(11), "pursuit": rh.policy_pursuit, "los_pn": lambda: rh.policy_los_pn(11),
           "mixed": lambda: rh.policy_mixed(11)}[policy]()
    out = dict(length=[], ret=[], min_distance=[], final_distance=[], cause=[], lock_fraction=[])
    for ep in range(n_ep):
        obs, _ = env.reset(seed=1000 + ep)
        ret, locks, t = 0.0, 0, 0
        while True:
            a = pol(t, obs[None, :])[0]
            obs, r, te, tr, info = env.step(a)
            ret += float(r); locks += int(bool(info["radar_detected"])); t += 1
            if te or tr:
                break
        out["length"].append(t); out["ret"].append(ret); out["min_distance"].append(float(info["min_distance"]))
        out["final_distance"].append(float(info["distance"])); out["cause"].append(cause_of(info, te, env))
        out["lock_fraction"].append(locks / t)
    arrs = {k: np.array(v) for k, v in out.items()}
    arrs["meta"] = np.array(json.dumps(dict(name=name, base=base, policy=policy, env_cfg=cfg, causes=CAUSES, training_step_count=tsc,
                                            numpy=np.__version__, generator="oracle/gen_stat_golden.py")))
    np.savez_compressed(os.path.join(GOLDEN_DIR, name + ".npz"), **arrs)
    print(name, "len", arrs["length"].mean(), "ret", arrs["ret"].mean(), "min_d", arrs["min_distance"].mean(),
          "causes", np.bincount(arrs["cause"], minlength=7))


if __name__ == "__main__":
    if not rh.reference_available():
        sys.exit("reference tree not found")
    for nm in (sys.argv[1:] or list(CASES)):
        generate(nm)
