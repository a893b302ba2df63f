"""The C oracle against (a) the golden fixtures generated from the unmodified reference and (b) the
reference itself when the reference tree is present (build container)."""
import numpy as np
import pytest

from common import GOLDEN_CASES, golden_setup, load_golden, replay_against_golden
from hlynr_intercept_b200 import config
from oracle import oracle, ref_harness


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_golden(name):
    g = load_golden(name)
    meta = g["meta"]
    P, cur = golden_setup(g)
    sim = oracle.OracleBatch(P, cur, meta["n_envs"], seed=meta["seed"], float64=meta["float64"])
    # The oracle emulates NumPy's rounding: positions, velocities, distances and rewards are bit-exact in the
    # float32 mode; only sin/cos/pow/atan2-derived values differ by an ulp.
    w = replay_against_golden(sim, g, rtol_state=5e-6, obs_atol=4e-6, reward_rtol=1e-4, reward_atol=1e-4,
                              margin_fn=lambda: sim.margin, margin_tol=1e-6, tti_atol=1e-4)
    assert w["dropped"] == 0
    assert sim.kalman_decoupling_error() == 0.0  # the 2x2-per-axis Kalman used on the GPU is exact


@pytest.mark.reference
@pytest.mark.parametrize("base,f64", [("cfg2", False), ("cfg4", True), ("cfg3", False)])
def test_oracle_matches_live_reference(base, f64):
    cfg = config.baseline_config(base)
    P, cur = config.resolve_config(cfg, warn_dead=False)
    n, T, seed = 4, 60, 99
    ref = ref_harness.RefBatch(cfg, n, seed=seed, float64=f64)
    sim = oracle.OracleBatch(P, cur, n, seed=seed, float64=f64)
    o_r, o_o = ref.reset(), sim.reset()
    np.testing.assert_allclose(o_o, o_r, atol=2e-6)
    pol = ref_harness.policy_random(3)
    for t in range(T):
        a = pol(t, o_r)
        o_r, r_r, te_r, tr_r, _, inf_r = ref.step(a)
        o_o, r_o, te_o, tr_o, _, inf_o = sim.step(a)
        assert (te_r == te_o).all() and (tr_r == tr_o).all() and (inf_r["flags"] == inf_o["flags"]).all()
        assert (inf_r["radar_quality"].astype(np.float32) == inf_o["radar_quality"]).all()
        np.testing.assert_allclose(o_o, o_r, atol=2e-5)
        np.testing.assert_allclose(r_o, r_r, rtol=2e-6, atol=2e-6)


def test_threads_equal_serial():
    P, cur = config.resolve_config(config.baseline_config("cfg4"), warn_dead=False)
    a = oracle.OracleBatch(P, cur, 64, seed=5)
    b = oracle.OracleBatch(P, cur, 64, seed=5, threads=4)
    oa, ob = a.reset(), b.reset()
    assert (oa == ob).all()
    rng = np.random.default_rng(0)
    for _ in range(20):
        act = rng.uniform(-1, 1, (64, 6)).astype(np.float32)
        ra, rb = a.step(act), b.step(act)
        assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all()
    assert a.stats() == b.stats()


def _reference_yaml_files():
    import glob
    import os

    root = os.path.join(ref_harness.REFERENCE_ROOT, "rl_system")
    files = sorted(glob.glob(os.path.join(root, "configs", "*.yaml")) + glob.glob(os.path.join(root, "configs", "*", "*.yaml")))
    return [os.path.relpath(f, root) for f in files + [os.path.join(root, "config.yaml")]]


@pytest.mark.reference
def test_oracle_matches_live_reference_on_every_reference_yaml():
    """Every YAML the reference ships (configs/*.yaml, configs/hrl/*.yaml, configs/scenarios/*.yaml, config.yaml), merged
    the way scripts/train_hrl_pretrain.py:270-338 merges it: the resolver accepts it and the oracle follows the unmodified
    reference for 500 ticks (integer outputs exact; the yaml_* fixtures carry each configuration through its terminations)."""
    from oracle import gen_golden

    files = _reference_yaml_files()
    assert len(files) >= 36
    n, T, seed = 2, 500, 77
    for rel in files:
        cfg = gen_golden.reference_yaml_env(rel)
        P, cur = config.resolve_config(cfg, warn_dead=False)
        ref = ref_harness.RefBatch(cfg, n, seed=seed)
        sim = oracle.OracleBatch(P, cur, n, seed=seed)
        pol = ref_harness.policy_true_guidance(ref, los_frame=cfg.get("observation_mode") == "los_frame")
        o_r, o_o = ref.reset(), sim.reset()
        np.testing.assert_allclose(o_o, o_r, atol=2e-6, err_msg=rel)
        for t in range(T):
            a = pol(t, o_r)
            o_r, r_r, te_r, tr_r, _, inf_r = ref.step(a)
            o_o, r_o, te_o, tr_o, _, inf_o = sim.step(a)
            assert (te_r == te_o).all() and (tr_r == tr_o).all() and (inf_r["flags"] == inf_o["flags"]).all(), (rel, t)
            np.testing.assert_allclose(o_o, o_r, atol=2e-5, err_msg=f"{rel} t={t}")
            np.testing.assert_allclose(r_o, r_r, rtol=1e-5, atol=1e-4, err_msg=f"{rel} t={t}")


@pytest.mark.reference
def test_oracle_matches_live_reference_on_mixed_feature_configs():
    """tests/sweep_configs.py: feature combinations no shipped YAML uses, the oracle next to the unmodified reference for 400
    ticks of smooth open-loop actions (integer outputs exact).  The GPU suite checks CUDA against the oracle on the same dicts."""
    import sweep_configs

    for k in range(sweep_configs.N_SWEEP):
        cfg = sweep_configs.sweep_config(k)
        P, cur = config.resolve_config(cfg, warn_dead=False)
        n, T = 2, 400
        ref = ref_harness.RefBatch(cfg, n, seed=55 + k)
        sim = oracle.OracleBatch(P, cur, n, seed=55 + k)
        pol = sweep_configs.sweep_policy(cfg, k)
        o_r, o_o = ref.reset(), sim.reset()
        np.testing.assert_allclose(o_o, o_r, atol=2e-6, err_msg=str(k))
        for t in range(T):
            a = pol(t, n)
            o_r, r_r, te_r, tr_r, _, inf_r = ref.step(a)
            o_o, r_o, te_o, tr_o, _, inf_o = sim.step(a)
            assert (te_r == te_o).all() and (tr_r == tr_o).all() and (inf_r["flags"] == inf_o["flags"]).all(), (k, t)
            assert (inf_r["radar_quality"].astype(np.float32) == inf_o["radar_quality"]).all(), (k, t)   # 0.0 while the onboard delay buffer fills
            np.testing.assert_allclose(o_o, o_r, atol=5e-5, err_msg=f"config {k} t={t}")
            np.testing.assert_allclose(r_o, r_r, rtol=1e-5, atol=1e-4, err_msg=f"config {k} t={t}")
