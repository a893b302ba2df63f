"""TEST INFRASTRUCTURE -- generates tests/golden/*.npz from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python -m oracle.gen_golden [case ...]

Every fixture is one run of oracle.ref_harness.RefBatch (the imported reference environment with
SB3-DummyVecEnv auto-reset and the site-keyed draws of include/hlynr_rng.h injected in place of its NumPy
generators).  The reference publishes no golden vectors of its own, so these fixtures are what
pins the oracle -- and through it the CUDA path -- to the reference.
"""
import json
import os
import sys

import numpy as np

from hlynr_intercept_b200 import config
from . import ref_harness as rh

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")

# merged curriculum as scripts/train_hrl_pretrain.py:337-338 would pass it (values of rl_system/config.yaml:74-111)
CONFIG_YAML_CURRICULUM = dict(
    enabled=True, initial_radius=100.0, final_radius=5.0, curriculum_steps=2000000,
    radar_curriculum=dict(enabled=True, initial_beam_width=120.0, final_beam_width=60.0,
                          beam_width_transition_start=5000000, beam_width_transition_end=8000000,
                          initial_detection_reliability=1.0, final_detection_reliability=0.8,
                          reliability_transition_start=1000000, reliability_transition_end=2000000,
                          initial_ground_reliability=1.0, final_ground_reliability=0.9,
                          ground_reliability_transition_start=1000000, ground_reliability_transition_end=2000000))


def case_config(base, extra=None):
    cfg = config.baseline_config(base)
    if extra:
        cfg.update(extra)
    return cfg


CASES = {
    # name: (base cfg, extra env keys, n_envs, T, float64, policy, training_step_count)
    "cfg1_f32_random_episode": ("cfg1", None, 1, 2000, False, "random", None),
    "cfg1_f64_random_episode": ("cfg1", None, 1, 2000, True, "random", None),
    "cfg2_f32_random": ("cfg2", None, 16, 128, False, "random", None),
    "cfg2_f64_random": ("cfg2", None, 16, 128, True, "random", None),
    "cfg3_f32_random": ("cfg3", None, 16, 128, False, "random", None),
    "cfg3_f64_random": ("cfg3", None, 16, 128, True, "random", None),
    "cfg3radar_f32_random": ("cfg3_radar", None, 16, 128, False, "random", None),
    "cfg4_f32_random": ("cfg4", None, 16, 128, False, "random", None),
    "cfg4_f64_random": ("cfg4", None, 16, 128, True, "random", None),
    "cfg4_f32_pursuit_long": ("cfg4", None, 8, 1500, False, "pursuit", None),
    "cfg4_f64_pursuit_long": ("cfg4", None, 4, 1500, True, "pursuit", None),
    "cfg3_f32_random_long": ("cfg3", None, 6, 1800, False, "random", None),
    "cfg2_f32_pursuit_long": ("cfg2", None, 6, 1500, False, "pursuit", None),
    "cfg4_f32_curriculum": ("cfg4", dict(curriculum=CONFIG_YAML_CURRICULUM), 8, 128, False, "pursuit", 1500000),
}


def reference_yaml_env(rel_path):
    """Env dict exactly as scripts/train_hrl_pretrain.py builds it: load_config (:270-296, child sections update the
    parent's) then environment + top-level curriculum + physics_enhancements (:333-338)."""
    import yaml

    path = os.path.join(rh.REFERENCE_ROOT, "rl_system", rel_path)
    with open(path) as f:
        cfg = yaml.safe_load(f)
    parent = os.path.normpath(os.path.join(os.path.dirname(path), cfg["parent_config"])) if "parent_config" in cfg else None
    if parent and os.path.exists(parent):   # a missing parent is skipped silently (:282; configs/hrl/eval_100m.yaml)
        with open(parent) as f:
            merged = yaml.safe_load(f)
        for key, val in cfg.items():
            if key == "parent_config":
                continue
            if isinstance(val, dict) and key in merged:
                merged[key].update(val)
            else:
                merged[key] = val
        cfg = merged
    env = dict(cfg.get("environment", {}))
    env["curriculum"] = cfg.get("curriculum", {})
    env["physics_enhancements"] = cfg.get("physics_enhancements", {})
    return env


# HRL specialist configs: precision-mode reward/termination, proximity fuze, spherical spawn,
# toward_missile launch, body_frame / los_frame observations and the LOS-frame action transform
CASES.update({
    "hrl_terminal360v1_f32_pursuit": ("yaml:configs/hrl/terminal_360_v1.yaml", None, 8, 1200, False, "pursuit", 400000),
    "hrl_terminal360v1_f32_zem": ("yaml:configs/hrl/terminal_360_v1.yaml", None, 8, 1200, False, "zem", 400000),
    "hrl_terminal360v1_f64_zem": ("yaml:configs/hrl/terminal_360_v1.yaml", None, 4, 1000, True, "zem", 400000),
    "hrl_rotinv_f32_mixed": ("yaml:configs/hrl/terminal_360_rotinv.yaml", None, 8, 900, False, "mixed", 1000000),
    "hrl_rotinv_f32_zem": ("yaml:configs/hrl/terminal_360_rotinv.yaml", None, 8, 1200, False, "zem", 1000000),
    "hrl_rotinv_f64_zem": ("yaml:configs/hrl/terminal_360_rotinv.yaml", None, 4, 800, True, "zem", 1000000),
    "hrl_terminal_los_f32_pn": ("yaml:configs/hrl/terminal_los.yaml", None, 8, 1200, False, "los_pn", 500000),
    "hrl_terminal_los_f32_zem": ("yaml:configs/hrl/terminal_los.yaml", None, 8, 1200, False, "zem_los", 500000),
    "hrl_terminal_los_f64_zem": ("yaml:configs/hrl/terminal_los.yaml", None, 4, 1000, True, "zem_los", 500000),
    "hrl_search_los_f32_random": ("yaml:configs/hrl/search_los.yaml", None, 8, 400, False, "random", None),
    # volley mode (inference.py:392-396): K missiles per env, priority = closest active missile
    "volley3_cfg4_f32_zem": ("cfg4", dict(volley_mode=True, volley_size=3), 8, 2300, False, "zem", None),
    "volley3_easy_f64_zem": ("cfg1", dict(volley_mode=True, volley_size=3), 4, 1500, True, "zem", None),
    "volley5_cfg2_f32_pursuit": ("cfg2", dict(volley_mode=True, volley_size=5), 6, 2100, False, "pursuit", None),
    "volley4_los_f32_zem": ("yaml:configs/hrl/terminal_los.yaml", dict(volley_mode=True, volley_size=4), 6, 1500, False, "zem_los", 500000),
    "volley1_cfg4_f32_zem": ("cfg4", dict(volley_mode=True, volley_size=1), 4, 800, False, "zem", None),
    "hrl_eval360los_f32_pn": ("yaml:configs/eval_360_los.yaml", None, 8, 1000, False, "los_pn", None),
})


# One short fixture per remaining reference YAML (every file under rl_system/configs plus config.yaml that has no case
# above), so that each configuration a reference user can pass is pinned: 2 envs, 1300 ticks (past the first
# terminations), true-state guidance in the frame the config's action space uses.  Left out because their merged env dict is
# identical to a listed one: hrl_base / selector_config / track_specialist (= config.yaml), terminal_specialist
# (= search_specialist), terminal_precision_v3 (= terminal_precision_v2).
YAML_SWEEP = [
    "config.yaml", "configs/eval_360_los_no_fuze.yaml", "configs/eval_360_proximity.yaml", "configs/eval_360_rotinv.yaml",
    "configs/eval_precision.yaml", "configs/eval_proximity_fuze.yaml", "configs/eval_terminal_standalone.yaml",
    "configs/hrl/eval_100m.yaml", "configs/hrl/hrl_curriculum.yaml",
    "configs/hrl/search_specialist.yaml", "configs/hrl/selector_los.yaml",
    "configs/hrl/terminal_2octant_rotinv.yaml", "configs/hrl/terminal_360_fresh.yaml", "configs/hrl/terminal_finetune.yaml",
    "configs/hrl/terminal_longrange_v1.yaml", "configs/hrl/terminal_precision.yaml", "configs/hrl/terminal_precision_v2.yaml",
    "configs/hrl/terminal_precision_v4.yaml", "configs/hrl/terminal_precision_v5.yaml",
    "configs/hrl/terminal_precision_v6.yaml", "configs/hrl/terminal_tight_30m.yaml",
    "configs/hrl/terminal_tight_50m.yaml", "configs/hrl/track_los.yaml",
    "configs/scenarios/easy.yaml", "configs/scenarios/medium.yaml", "configs/scenarios/hard.yaml",
]
for _rel in YAML_SWEEP:
    _stem = os.path.splitext(os.path.basename(_rel))[0]
    _scen = "scenario_" if "/scenarios/" in _rel else ""
    CASES["yaml_" + _scen + _stem] = ("yaml:" + _rel, None, 2, 1300, False, "zem_auto", None)


def make_policy(policy, ref):
    return {"random": lambda: rh.policy_random(7), "pursuit": rh.policy_pursuit, "mixed": lambda: rh.policy_mixed(11),
            "los_pn": lambda: rh.policy_los_pn(13), "zem": lambda: rh.policy_true_guidance(ref),
            "zem_los": lambda: rh.policy_true_guidance(ref, los_frame=True),
            "zem_auto": lambda: rh.policy_true_guidance(ref, los_frame=ref.envs[0].config.get("observation_mode") == "los_frame"),
            }[policy]()


def generate(name):
    base, extra, n, T, f64, policy, tsc = CASES[name]
    cfg = reference_yaml_env(base[5:]) if base.startswith("yaml:") else case_config(base, None)
    if extra:
        cfg.update(extra)
    seed = 1234
    ref = rh.RefBatch(cfg, n, seed=seed, float64=f64, training_step_count=tsc)
    pol = make_policy(policy, ref)
    obs0 = ref.reset()
    obs = obs0
    rec = dict(actions=[], obs=[], reward=[], terminated=[], truncated=[], terminal_obs=[], distance=[],
               min_distance=[], fuel_remaining=[], fuel_used=[], steps=[], flags=[], interceptor_pos=[],
               missile_pos=[], episode_return=[], episode_length=[])
    volley_keys = ("missiles_intercepted", "missiles_remaining", "missile_min_distances") if cfg.get("volley_mode") else ()
    for k in volley_keys:
        rec[k] = []
    stop_after_first_done = name.startswith("cfg1")
    for t in range(T):
        a = pol(t, obs)
        obs, r, te, tr, tobs, info = ref.step(a)
        rec["actions"].append(a)
        rec["obs"].append(obs)
        rec["reward"].append(r)
        rec["terminated"].append(te)
        rec["truncated"].append(tr)
        rec["terminal_obs"].append(tobs)
        for k in ("distance", "min_distance", "fuel_remaining", "fuel_used", "steps", "flags", "interceptor_pos",
                  "missile_pos", "episode_return", "episode_length") + volley_keys:
            rec[k].append(info[k])
        if stop_after_first_done and (te | tr).any():
            break
    out = {k: np.stack(v) for k, v in rec.items()}
    out["actions"] = out["actions"].astype(np.float32)
    out["obs"] = out["obs"].astype(np.float32)
    done = (out["terminated"] | out["truncated"]).astype(bool)
    # terminal observations only where an episode ended (sparse)
    idx = np.argwhere(done)
    out["terminal_idx"] = idx.astype(np.int32)
    out["terminal_obs"] = out["terminal_obs"][done].astype(np.float32)
    for k in ("distance", "min_distance", "fuel_remaining", "fuel_used", "interceptor_pos", "missile_pos",
              "episode_return"):
        out[k] = out[k].astype(np.float64 if f64 else np.float32)
    out["obs0"] = obs0
    st = ref.export_state()
    for k, v in st.items():
        out["final_" + k] = v
    meta = dict(name=name, base=base, env_cfg=cfg, n_envs=n, steps=int(out["obs"].shape[0]), float64=f64, policy=policy,
                training_step_count=tsc, seed=seed, curriculum=ref.curriculum(), numpy=np.__version__,
                generator="oracle/gen_golden.py", reference="RomanSlack/Hlynr_Intercept rl_system/environment.py")
    out["meta"] = np.array(json.dumps(meta))
    os.makedirs(GOLDEN_DIR, exist_ok=True)
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    np.savez_compressed(path, **out)
    n_done = int(done.sum())
    print(f"{name}: T={out['obs'].shape[0]} n={n} episodes_finished={n_done} "
          f"intercepts={int(((out['flags'] & 1) > 0)[done].sum())} -> {os.path.getsize(path) / 1e3:.0f} kB")


def load(name):
    path = os.path.join(GOLDEN_DIR, name + ".npz")
    z = np.load(path, allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


if __name__ == "__main__":
    if not rh.reference_available():
        sys.exit("reference tree not found at " + rh.REFERENCE_ROOT)
    names = sys.argv[1:] or list(CASES)
    for nm in names:
        generate(nm)
