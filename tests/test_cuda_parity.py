"""GPU parity: the CUDA path (through the C ABI) against the golden fixtures generated from the unmodified
reference and against the C oracle on identical seeds, actions and injected draws.

Tolerances are the north_star's: integers / booleans bit-exact; floats rtol 1e-3 (fp32 build vs native
reference) and rtol 1e-5 (fp64 build vs up-cast float64 reference), over 1-step and >=100-step horizons.
Envs whose decision margin (distance to a threshold, reported by the oracle) is below the tolerance AND that
actually disagree are excluded from the exact comparison from that tick on, and counted.

Every test appends its worst observed errors to gpurun_out/parity_report.jsonl (tests/common.ParityLog);
tools/parity_report.py folds a GPU run's records into the committed profiles/parity_report.json.
"""
import numpy as np
import pytest

import sweep_configs
from common import (GOLDEN_CASES, STATE_FLOAT_FIELDS, STATE_INT_FIELDS, CudaBatch, Lockstep, ParityLog, TOL, golden_setup,
                    load_golden, replay_against_golden)
from hlynr_intercept_b200 import config
from oracle import draws, oracle

pytestmark = pytest.mark.gpu


def test_philox_on_device_matches_contract():
    P, cur = config.resolve_config(config.baseline_config("cfg2"), warn_dead=False)
    sim = CudaBatch(P, cur, 32, seed=0x1234567890AB).sim
    for env, ep, st, blk in [(0, 0, 0, 0), (5, 3, 17, 2), (2 ** 33 + 7, 9, 1999, 4), (123456, 0, 1, 16)]:
        raw, uni, nrm = sim.debug_draws(env, ep, st, blk)
        want = draws.block(0x1234567890AB, env, ep, st, blk)
        assert (raw == want).all()
        assert (uni == draws.u01(want)).all()
        np.testing.assert_allclose(nrm, draws.normals(want), rtol=2e-6, atol=2e-7)


TRAJ_CASES = [c for c in GOLDEN_CASES if not c.startswith("stat_")]


@pytest.mark.parametrize("name", TRAJ_CASES)
def test_cuda_matches_reference_golden(name):
    g = load_golden(name)
    meta = g["meta"]
    P, cur = golden_setup(g)
    f64 = meta["float64"]
    cuda = CudaBatch(P, cur, meta["n_envs"], seed=meta["seed"], float64=f64)
    shadow = oracle.OracleBatch(P, cur, meta["n_envs"], seed=meta["seed"], float64=f64)
    sim = Lockstep(cuda, shadow)
    tol = TOL[f64]
    log = ParityLog(f"golden/{name}", tol["rtol_state"], f64=bool(f64), n_envs=int(meta["n_envs"]), ticks=int(g["obs"].shape[0]))
    w = replay_against_golden(sim, g, margin_fn=lambda: sim.margin, log=log, **tol)
    log.write()
    assert w["dropped"] == 0, w


class PairChecker:
    """CUDA next to the oracle on the same actions, tick by tick: integer outputs exact, floats within the build's tolerance;
    envs whose oracle decision margin is below the tolerance AND that actually disagree are dropped and counted."""

    def __init__(self, n, f64, log, loose=(9, 10, 11, 13, 16)):
        self.tol = dict(TOL[f64])
        self.f64 = bool(f64)
        self.loose = list(loose)
        self.alive = np.ones(n, bool)
        self.twin_ok = np.ones(n, bool)   # the twin follows the same episode schedule as long as its done flags agree
        self.log = log

    def tick(self, t, cuda_out, orc_out, margin, twin_out=None):
        tol, loose = self.tol, self.loose
        oc, rc, tec, trc, ic = cuda_out
        oo, ro, teo, tro, io = orc_out
        low = margin < tol["margin_tol"]
        x13 = np.maximum(1.0, 4.0 * (1.0 - oo[:, 13]) ** 2)   # obs[13] sensitivity to the closing speed (tests/common.py)
        d_all = np.abs(oc.astype(np.float64) - oo)
        d_all[:, 13] /= x13
        if 2 in loose:   # los_frame: channels 2, 3 are LOS rates = transverse velocity / estimated range (obs[0] * max_range):
            d_all[:, 2:4] /= np.maximum(1.0, 0.01 / np.maximum(oo[:, 0], 1e-9))[:, None]   # below 100 m the tolerance grows like 1 / range
        for ch in (9, 11):   # roll / yaw over pi: -1 and +1 are the same angle
            d_all[:, ch] = np.minimum(d_all[:, ch], 2.0 - d_all[:, ch])
        if twin_out is not None:
            # the oracle in the OTHER precision on the same draws and actions: where the reference's own float32 and float64
            # evaluations of an observation element disagree by delta (LOS rates right after the Kalman initialisation or at
            # short range, cosines of nearly-zero vectors), no implementation can be pinned tighter than a few delta
            ot, tet, trt = twin_out
            self.twin_ok &= (tet == teo) & (trt == tro)
            slack = np.where(self.twin_ok[:, None], 3.0 * np.abs(ot - oo), 0.0)
            d_all = np.maximum(d_all - slack, 0.0)
        differs = (tec != teo) | (trc != tro) | (ic["flags"] != io["flags"]) | (d_all.max(axis=1) > tol["tti_atol"])
        newly = self.alive & low & differs
        self.alive &= ~newly
        self.log.dropped += int(newly.sum())
        a = self.alive
        assert (tec[a] == teo[a]).all() and (trc[a] == tro[a]).all(), f"done mismatch at t={t}"
        assert (ic["flags"][a] == io["flags"][a]).all(), f"flag mismatch at t={t}"
        assert (ic["steps"][a] == io["steps"][a]).all()
        self.log.obs(oc[a], oo[a])
        self.log.field("reward", rc[a], ro[a])
        for k in ("distance", "min_distance", "fuel_remaining", "fuel_used"):
            self.log.field(k, ic[k][a], io[k][a])
        d = d_all[a]
        worst_loose = d[:, loose].max() if d.size else 0.0
        where_loose = np.unravel_index(d[:, loose].argmax(), d[:, loose].shape) if d.size else None
        d[:, loose] = 0
        assert d.size == 0 or d.max() <= tol["obs_atol"], f"obs mismatch t={t}: {d.max()} idx {np.unravel_index(d.argmax(), d.shape)}"
        assert worst_loose <= tol["tti_atol"], (f"ill-conditioned obs mismatch t={t}: {worst_loose} "
                                                 f"(alive env #{where_loose[0]}, channel {loose[where_loose[1]]}, oracle row {oo[a][where_loose[0]][:8]}, "
                                                 f"cuda row {oc[a][where_loose[0]][:8]}, distance {io['distance'][a][where_loose[0]]})")
        err = np.abs(rc[a] - ro[a])
        assert (err <= tol["reward_atol"] + tol["reward_rtol"] * np.abs(ro[a])).all(), f"reward mismatch t={t}: {err.max()}"
        for k in ("distance", "min_distance", "fuel_remaining", "fuel_used"):
            np.testing.assert_allclose(ic[k][a], io[k][a], rtol=tol["rtol_state"], atol=tol["reward_atol"])

    def final_state(self, sc, so):
        a = self.alive
        for k in STATE_INT_FIELDS:
            assert (sc[k][a] == so[k][a]).all(), k
        if self.f64:   # the fp64 build tracks the dtype of the reference's Kalman state array (float32 -> float64 switch)
            assert (sc["kf_f64"][a] == so["kf_f64"][a]).all(), "kf_f64"
        for k in STATE_FLOAT_FIELDS:
            ref = so[k][a]
            self.log.field("final_" + k, sc[k][a], ref)
            np.testing.assert_allclose(sc[k][a], ref, rtol=self.tol["rtol_state"],
                                       atol=self.tol["rtol_state"] * (np.abs(ref).max() + 1e-6), err_msg=k)


def _compare_cuda_with_oracle(test, P, cur, n, T, seed, f64, action_fn=None, max_dropped=0, conditioning=False,
                              loose=(9, 10, 11, 13, 16)):
    cuda = CudaBatch(P, cur, n, seed=seed, float64=f64)
    orc = oracle.OracleBatch(P, cur, n, seed=seed, float64=f64, threads=8)
    twin = oracle.OracleBatch(P, cur, n, seed=seed, float64=not f64, threads=8) if conditioning else None
    log = ParityLog(test, TOL[f64]["rtol_state"], f64=bool(f64), n_envs=n, ticks=T)
    chk = PairChecker(n, f64, log, loose)
    o_c, o_o = cuda.reset(), orc.reset()
    if twin is not None:
        twin.reset()
    np.testing.assert_allclose(o_c, o_o, rtol=0, atol=chk.tol["obs_atol"])
    env_ids = np.arange(n)
    episode = np.zeros(n, np.int64)
    steps = np.zeros(n, np.int64)
    for t in range(T):
        act = draws.random_actions(seed, env_ids, episode, steps + 1) if action_fn is None else action_fn(t, n)
        oc, rc, tec, trc, _, ic = cuda.step(act)
        oo, ro, teo, tro, _, io = orc.step(act)
        tw = None
        if twin is not None:
            ot, _, tet, trt, _, _ = twin.step(act)
            tw = (ot, tet, trt)
        chk.tick(t, (oc, rc, tec, trc, ic), (oo, ro, teo, tro, io), orc.margin, tw)
        done = (teo | tro).astype(bool)
        episode += done
        steps = np.where(done, 0, steps + 1)
    chk.final_state(cuda.export_state(), orc.export_state())
    log.write(episodes=int(episode.sum()))
    assert log.dropped <= max_dropped, f"low-margin exclusions: {log.dropped}"
    return int(episode.sum())


@pytest.mark.parametrize("base", ["cfg2", "cfg3", "cfg4"])
@pytest.mark.parametrize("f64", [False, True])
def test_cuda_matches_oracle_4096_envs_100_steps(base, f64):
    """BASELINE cfg2 size (4096 envs, here 4100 to exercise a partial tile): CUDA vs oracle, 1-step and 100-step horizon."""
    P, cur = config.resolve_config(config.baseline_config(base), warn_dead=False)
    _compare_cuda_with_oracle(f"oracle4100/{base}/{'f64' if f64 else 'f32'}", P, cur, 4100, 100, 4242, f64, max_dropped=0)


@pytest.mark.parametrize("f64", [False, True])
@pytest.mark.parametrize("k", range(sweep_configs.N_SWEEP))
def test_cuda_matches_oracle_on_mixed_feature_configs(k, f64):
    """Feature combinations no shipped YAML uses (tests/sweep_configs.py: every physics v2.0 sub-switch on its own, odd
    delays, no ground radar, spherical spawns, precision mode / fuze / volley / observation modes on top of domain
    randomization): 1030 envs x 450 ticks of smooth open-loop actions, through resets.  The oracle is pinned to the
    unmodified reference on the very same dicts by tests/test_oracle.py::test_oracle_matches_live_reference_on_mixed_feature_configs.
    The fp64 build runs at the north star's 1e-5 (it reproduces the reference's float32 -> float64 switch of the Kalman
    state); in the fp32 build the los_frame channels 2-5 (LOS rates and direction cosines of ESTIMATED velocities, which
    amplify float32 rounding of the track filter by its ~1/dt gains and by 1/range) share the ill-conditioned-channel bucket."""
    if f64 and k % 3:
        pytest.skip("fp64 build: every third configuration")
    cfg = sweep_configs.sweep_config(k)
    P, cur = config.resolve_config(cfg, warn_dead=False)
    loose = (9, 10, 11, 13, 16) + ((2, 3, 4, 5) if (cfg.get("observation_mode") == "los_frame" and not f64) else ())
    _compare_cuda_with_oracle(f"sweep/{k}/{'f64' if f64 else 'f32'}", P, cur, 1030, 450, 900 + k, f64,
                              action_fn=sweep_configs.sweep_policy(cfg, k), max_dropped=2, conditioning=True, loose=loose)


@pytest.mark.parametrize("base,n,f64", [("cfg4", 1 << 20, False), ("cfg3", 262144, False), ("cfg4", (1 << 20) + 77, False),
                                        ("cfg3", 262144 - 51, True)])
def test_full_size_windows_match_oracle(base, n, f64):
    """BASELINE sizes (cfg4 at 2^20 envs, cfg3 at 262144; plus ragged variants of both, one in the fp64 build): the whole batch
    steps 120 ticks on the GPU and three windows of it -- the first tiles, tiles in the middle, and the last (ragged) tiles --
    are compared tick by tick with the oracle stepping the same GLOBAL env ids (OracleBatch(env_id_offset=...))."""
    import torch

    P, cur = config.resolve_config(config.baseline_config(base), warn_dead=False)
    seed, T, w = 20240 + (n & 0xff), 120, 320
    sim = CudaBatch(P, cur, n, seed=seed, float64=f64).sim
    starts = [0, (n // 2 // 128) * 128 + 64, n - w]
    idx = torch.cat([torch.arange(s, s + w) for s in starts]).cuda()
    orcs = [oracle.OracleBatch(P, cur, w, seed=seed, env_id_offset=s, float64=f64, threads=4) for s in starts]
    log = ParityLog(f"windows/{base}/{n}/{'f64' if f64 else 'f32'}", TOL[f64]["rtol_state"], f64=bool(f64), n_envs=n, ticks=T,
                    windows=[[s, s + w] for s in starts])
    chk = PairChecker(3 * w, f64, log)
    oc = sim.reset().index_select(0, idx).cpu().numpy()
    oo = np.concatenate([o.reset() for o in orcs])
    np.testing.assert_allclose(oc, oo, rtol=0, atol=chk.tol["obs_atol"])
    g = torch.Generator(device="cuda")
    g.manual_seed(seed)
    for t in range(T):
        act = (torch.rand(n, 6, device="cuda", generator=g) * 2 - 1).contiguous()
        obs, rew, te, tr, _, info = sim.step(act, want_info=True, want_terminal_obs=False)
        sel = lambda x: x.index_select(0, idx).cpu().numpy()  # noqa: E731
        ic = {k: sel(info[k]) for k in ("flags", "steps", "distance", "min_distance", "fuel_remaining", "fuel_used")}
        cuda_out = (sel(obs), sel(rew).astype(np.float64), sel(te), sel(tr), ic)
        a_h = act.index_select(0, idx).cpu().numpy()
        outs = [o.step(a_h[k * w:(k + 1) * w]) for k, o in enumerate(orcs)]
        cat = lambda j: np.concatenate([o[j] for o in outs])  # noqa: E731
        io = {k: np.concatenate([o[5][k] for o in outs]) for k in ic}
        margin = np.concatenate([o.margin for o in orcs])
        chk.tick(t, cuda_out, (cat(0), cat(1), cat(2), cat(3), io), margin)
    sc = [sim.export_state(s, w) for s in starts]
    so = [o.export_state() for o in orcs]
    chk.final_state({k: np.concatenate([x[k] for x in sc]) for k in sc[0]}, {k: np.concatenate([x[k] for x in so]) for k in so[0]})
    log.write()
    assert log.dropped == 0, log.dropped
    # size-independent invariants over the WHOLE batch after the 120 ticks
    assert torch.isfinite(obs).all() and obs.min() >= -2.0 and obs.max() <= 1.0
    assert sim.stats()["env_steps"] == T * n


def test_sharding_invariance_and_rollout_equivalence():
    """Global env ids key the RNG: two half shards == one full batch, bit for bit; and the fused k-step rollout
    kernel == k single-step launches, bit for bit."""
    import torch

    P, cur = config.resolve_config(config.baseline_config("cfg4"), warn_dead=False)
    n, seed, K = 512, 77, 40
    full = CudaBatch(P, cur, n, seed=seed)
    lo = CudaBatch(P, cur, n // 2, seed=seed, env_id_offset=0)
    hi = CudaBatch(P, cur, n // 2, seed=seed, env_id_offset=n // 2)
    fused = CudaBatch(P, cur, n, seed=seed)
    o_full = full.reset()
    assert (np.concatenate([lo.reset(), hi.reset()]) == o_full).all()
    assert (fused.reset() == o_full).all()
    rng = np.random.default_rng(1)
    acts = rng.uniform(-1, 1, (K, n, 6)).astype(np.float32)
    rsum = np.zeros(n, np.float64)
    for k in range(K):
        of, rf, tef, trf, _, _ = full.step(acts[k])
        ol, rl, tel, trl, _, _ = lo.step(acts[k][: n // 2])
        oh, rh, teh, trh, _, _ = hi.step(acts[k][n // 2:])
        assert (np.concatenate([ol, oh]) == of).all() and (np.concatenate([rl, rh]) == rf).all()
        assert (np.concatenate([tel, teh]) == tef).all()
        rsum += rf
    obs_k, rs, dc = fused.sim.rollout(K, torch.as_tensor(acts).cuda())
    torch.cuda.synchronize()
    assert (obs_k.cpu().numpy() == of).all()
    np.testing.assert_allclose(rs.cpu().numpy(), rsum, rtol=1e-5, atol=1e-3)
    sa, sb = full.export_state(), fused.export_state()
    for k in STATE_INT_FIELDS + STATE_FLOAT_FIELDS:
        assert (sa[k] == sb[k]).all(), k


def test_random_policy_rollout_matches_host_actions():
    """hlynr_rollout(actions=NULL) uses the BLK_ACT draws: same trajectory as stepping with draws.random_actions."""
    P, cur = config.resolve_config(config.baseline_config("cfg2"), warn_dead=False)
    n, seed, K = 256, 31, 25
    a = CudaBatch(P, cur, n, seed=seed)
    b = CudaBatch(P, cur, n, seed=seed)
    a.reset(); b.reset()
    ids = np.arange(n)
    for k in range(K):
        oa = a.step(draws.random_actions(seed, ids, 0, k + 1))[0]
    ob, _, _ = b.sim.rollout(K, None)
    assert (ob.cpu().numpy() == oa).all()


def test_episode_statistics_match_oracle():
    g = load_golden("cfg4_f32_pursuit_long")
    P, cur = golden_setup(g)
    n = g["meta"]["n_envs"]
    cuda = CudaBatch(P, cur, n, seed=g["meta"]["seed"])
    orc = oracle.OracleBatch(P, cur, n, seed=g["meta"]["seed"])
    cuda.reset(); orc.reset()
    for t in range(g["obs"].shape[0]):
        cuda.step(g["actions"][t]); orc.step(g["actions"][t])
    sc, so = cuda.stats(), orc.stats()
    for k in ("episodes", "successes", "hit_target", "interceptor_crash", "fuel_out", "missile_ground", "worsening",
              "timeouts", "env_steps", "onboard_locks", "length_sum"):
        assert sc[k] == so[k], (k, sc[k], so[k])
    for k in ("return_sum", "min_distance_sum", "final_distance_sum"):
        np.testing.assert_allclose(sc[k], so[k], rtol=1e-3)
    assert sc["episodes"] == int((g["terminated"] | g["truncated"]).sum())


def test_full_size_invariants_1m_envs():
    """BASELINE cfg4 size (2^20 envs): size-independent properties of the fused rollout."""
    import torch

    n = 1 << 20
    P, cur = config.resolve_config(config.baseline_config("cfg4"), warn_dead=False)
    sim = CudaBatch(P, cur, n, seed=9).sim
    obs = sim.reset()
    assert torch.isfinite(obs).all() and obs.min() >= -2.0 and obs.max() <= 1.0
    assert (obs[:, 12] == 1.0).all()  # full fuel
    obs2, rsum, dcount = sim.rollout(64, None)
    torch.cuda.synchronize()
    assert torch.isfinite(obs2).all() and obs2.min() >= -2.0 and obs2.max() <= 1.0
    assert torch.isfinite(rsum).all()
    st = sim.export_state(0, 4096)
    q = st["quat"]
    np.testing.assert_allclose((q * q).sum(axis=1), 1.0, atol=1e-5)
    assert (st["steps"] == 64).all() and (st["episode"] == 0).all()
    assert (st["fuel"] <= 100.0).all() and (st["fuel"] > 90.0).all()
    s = sim.stats()
    assert s["env_steps"] == 64 * n
    # a reset with an all-zero mask changes nothing
    small = CudaBatch(P, cur, 4096, seed=9).sim
    first = small.reset().cpu().clone()
    again = small.reset(torch.zeros(4096, dtype=torch.uint8)).cpu()
    assert (again == first).all()
    assert (small.export_state()["episode"] == 0).all()


@pytest.mark.parametrize("f64", [False, True])
def test_manual_reset_flow_without_auto_reset(f64):
    """The Gymnasium single-env flow (inference.py:500-520): step(auto_reset=False) returns the terminal observation as
    obs and leaves the env alone; the caller resets the finished envs with a mask.  CUDA vs oracle, flags exact."""
    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 45
    P, cur = config.resolve_config(cfg, warn_dead=False)
    n = 333
    cuda = CudaBatch(P, cur, n, seed=21, float64=f64)
    orc = oracle.OracleBatch(P, cur, n, seed=21, float64=f64)
    oc, oo = cuda.reset(), orc.reset()
    np.testing.assert_allclose(oc, oo, atol=1e-3)
    rng = np.random.default_rng(4)
    tol = TOL[f64]
    resets = 0
    for t in range(140):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        c = cuda.step(a, auto_reset=False)
        o = orc.step(a, auto_reset=False)
        assert (c[2] == o[2]).all() and (c[3] == o[3]).all() and (c[5]["flags"] == o[5]["flags"]).all(), t
        assert (c[5]["steps"] == o[5]["steps"]).all()
        d = np.abs(c[0] - o[0]); d[:, [9, 10, 11, 13, 16]] = 0
        assert d.max() <= tol["obs_atol"], (t, d.max())
        np.testing.assert_allclose(c[1], o[1], rtol=tol["reward_rtol"], atol=tol["reward_atol"])
        done = (c[2] | c[3]).astype(bool)
        # reset only every other finished env now: the others keep stepping past their end, as the reference env allows
        mask = (done & (np.arange(n) % 2 == 0)) | (o[5]["steps"] > 60)
        if mask.any():
            rc, ro = cuda.reset(mask.astype(np.uint8)), orc.reset(mask.astype(np.uint8))
            np.testing.assert_allclose(rc[mask], ro[mask], atol=1e-3)
            resets += int(mask.sum())
    assert resets > n
    sc, so = cuda.export_state(), orc.export_state()
    for k in STATE_INT_FIELDS:
        assert (sc[k] == so[k]).all(), k
    for k in ("ipos", "ivel", "mpos", "mvel", "fuel"):
        np.testing.assert_allclose(sc[k], so[k], rtol=tol["rtol_state"], atol=tol["rtol_state"] * 10)


@pytest.mark.parametrize("base", ["cfg4", "cfg3"])
def test_compact_layout_is_bit_identical(monkeypatch, base):
    """The compact 10-plane layout of the cfg4 feature set (fp32 build: the counters ride in the constant words r6.w / f1.w and
    the i0 plane is never touched, hlynr_device.cuh load_env) against the 11-plane layout (HLYNR_NO_COMPACT=1 at create time) on
    the same seed and actions: every output of every tick, the info arrays, the terminal observations and the exported env
    state bit for bit -- early in the episode and at the steady state with auto-resets in every tick, ragged last tile included;
    plus an export -> import round trip into a fresh compact handle."""
    import torch

    # cfg3: the second compact layout (domain randomization: the drag peak rides in i0.y and the f3 plane is never touched)
    P, cur = config.resolve_config(config.baseline_config(base), warn_dead=False)
    n = 30000 + 13
    a = CudaBatch(P, cur, n, seed=777).sim
    monkeypatch.setenv("HLYNR_NO_COMPACT", "1")
    b = CudaBatch(P, cur, n, seed=777).sim
    monkeypatch.delenv("HLYNR_NO_COMPACT")
    assert torch.equal(a.reset(), b.reset())
    g = torch.Generator(device="cuda")
    g.manual_seed(5)
    resets = 0
    for phase, ticks in (("early", 30), ("steady", 120)):
        if phase == "steady":
            a.rollout(1700, None, want_obs=False)
            b.rollout(1700, None, want_obs=False)
        for t in range(ticks):
            act = (torch.rand(n, 6, device="cuda", generator=g) * 2.4 - 1.2).contiguous()
            oa = a.step(act, want_info=True)
            ob = b.step(act, want_info=True)
            for x, y, name in zip(oa[:4], ob[:4], ("obs", "reward", "terminated", "truncated")):
                assert torch.equal(x, y), (phase, t, name)
            done = (oa[2] | oa[3]).bool()
            resets += int(done.sum())
            assert torch.equal(oa[4][done], ob[4][done]), (phase, t, "terminal_obs")
            for k in oa[5]:
                assert torch.equal(oa[5][k], ob[5][k]), (phase, t, k)
    assert resets > 0
    sa, sb = a.export_state(), b.export_state()
    for k in sa:
        assert np.array_equal(sa[k], sb[k], equal_nan=True), k
    c = CudaBatch(P, cur, n, seed=777).sim   # a fresh compact handle continues from the exported state exactly
    c.reset()
    c.import_state(sa)
    sc = c.export_state()
    for k in sa:
        assert np.array_equal(sa[k], sc[k], equal_nan=True), k
