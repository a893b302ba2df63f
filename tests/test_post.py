"""On-device VecFrameStack + VecNormalize (include/hlynr_post.h) against the numpy restatement of
the Stable-Baselines3 wrappers (oracle/sb3_post.py).  SB3 itself is not importable here: parity unpinned, the
restatement follows SB3's published source."""
import numpy as np
import pytest

from hlynr_intercept_b200 import config
from oracle import sb3_post


def test_sb3_restatement_known_answers():
    """CPU: the restated wrappers on hand-computable inputs (RunningMeanStd merge, stack roll, done handling)."""
    rms = sb3_post.RunningMeanStd(shape=(2,))
    a = np.array([[1.0, 2.0], [3.0, 6.0]])
    rms.update(a)
    # count = 1e-4 + 2; mean = 0 + batch_mean * 2 / tot
    tot = 2 + 1e-4
    np.testing.assert_allclose(rms.mean, np.array([2.0, 4.0]) * 2 / tot, rtol=1e-12)
    m2 = 1.0 * 1e-4 + np.array([1.0, 4.0]) * 2 + np.array([4.0, 16.0]) * 1e-4 * 2 / tot
    np.testing.assert_allclose(rms.var, m2 / tot, rtol=1e-12)
    st = sb3_post.StackedObservations(2, 3, 2)
    o = st.reset(np.array([[1, 2], [3, 4]], np.float32))
    assert (o == np.array([[0, 0, 0, 0, 1, 2], [0, 0, 0, 0, 3, 4]], np.float32)).all()
    o, _ = st.update(np.array([[5, 6], [7, 8]], np.float32), np.array([False, False]), {})
    assert (o[0] == [0, 0, 1, 2, 5, 6]).all()
    infos = {1: {"terminal_observation": np.array([9, 9], np.float32)}}
    o, infos = st.update(np.array([[10, 11], [12, 13]], np.float32), np.array([False, True]), infos)
    assert (o[0] == [1, 2, 5, 6, 10, 11]).all() and (o[1] == [0, 0, 0, 0, 12, 13]).all()
    assert (infos[1]["terminal_observation"] == [3, 4, 7, 8, 9, 9]).all()   # previous stack shifted + terminal obs


def _staggered(sim, rng, max_steps):
    """Gives every env a different elapsed step count so truncations (max_steps) are spread over time."""
    st = sim.export_state()
    st["steps"] = rng.integers(0, max_steps - 3, size=sim.n).astype(np.int32)
    return st


@pytest.mark.gpu
@pytest.mark.parametrize("n_stack", [4, 1, 3])
def test_pipeline_matches_sb3_restatement(n_stack):
    import torch

    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.sim import HlynrSim

    n, max_steps, T = 3001, 40, 130
    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = max_steps
    ref = HlynrSim(cfg, n_envs=n, seed=5, warn_dead=False)
    sim = HlynrSim(cfg, n_envs=n, seed=5, warn_dead=False)
    pipe = HlynrObsPipeline(sim, n_stack=n_stack, training=True)
    out0 = pipe.reset()
    obs0 = ref.reset().cpu().numpy()
    rng = np.random.default_rng(1)
    state = _staggered(ref, rng, max_steps)
    ref.import_state(state); sim.import_state(state)
    stack = sb3_post.StackedObservations(n, n_stack, 26)
    vn = sb3_post.VecNormalize(n, (26 * n_stack,), gamma=0.99)
    want0 = vn.reset(stack.reset(obs0).astype(np.float64))
    np.testing.assert_allclose(out0.cpu().numpy(), want0, atol=2e-6)
    total_done = 0
    for t in range(T):
        if t == 70:  # all envs truncate together from here on (every record slot is used) ...
            pipe.training = False
            vn.training = False
        if t == 100:  # ... and statistics frozen for a while (inference.py:468), then unfrozen again
            pipe.training = True
            vn.training = True
        a = torch.as_tensor(rng.uniform(-1, 1, (n, 6)).astype(np.float32)).cuda()
        o, r, te, tr, tobs, _ = ref.step(a)
        o, r, tobs = o.cpu().numpy(), r.cpu().numpy(), tobs.cpu().numpy()
        dones = (te | tr).cpu().numpy().astype(bool)
        infos = {int(i): {"terminal_observation": tobs[i].copy()} for i in np.nonzero(dones)[0]}
        stacked, infos = stack.update(o, dones, infos)
        want, _, _, infos = vn.step(stacked.astype(np.float64), r.astype(np.float64), dones, infos)
        out, rew, te2, tr2, (records, counter, terminal) = pipe.step(a)
        assert (rew.cpu().numpy() == r).all() and ((te2 | tr2).cpu().numpy().astype(bool) == dones).all()
        np.testing.assert_allclose(out.cpu().numpy(), want, atol=5e-6, err_msg=f"tick {t}")
        np.testing.assert_array_equal(pipe.get_original_obs().cpu().numpy(), stacked)
        rec = pipe.done_records()
        assert sorted(rec["env"].tolist()) == sorted(infos.keys())
        term = terminal.cpu().numpy()
        for k, e in enumerate(rec["env"].tolist()):
            np.testing.assert_allclose(term[k], infos[e]["terminal_observation"], atol=5e-6)  # stacked by VecFrameStack, normalised by VecNormalize
        total_done += len(rec)
        s = pipe.get_stats()
        np.testing.assert_allclose(s["mean"], vn.obs_rms.mean, rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(s["var"], vn.obs_rms.var, rtol=1e-8, atol=1e-12)
        assert abs(s["count"] - vn.obs_rms.count) < 1e-6
        np.testing.assert_allclose([s["ret_mean"], s["ret_var"]], [vn.ret_rms.mean, vn.ret_rms.var], rtol=1e-9)
        if t % 16 == 0:
            assert pipe.check_sums() < 1e-6    # incrementally maintained per-lag sums vs a full recomputation
    assert total_done > 3 * n
    x = torch.as_tensor(rng.uniform(-2, 1, (7, 26 * n_stack)).astype(np.float32)).cuda()
    np.testing.assert_allclose(pipe.normalize_obs(x).cpu().numpy(), vn.normalize_obs(x.cpu().numpy().astype(np.float64)), atol=5e-6)
    # statistics round trip (the content of vec_normalize.pkl)
    s = pipe.get_stats()
    pipe.set_stats(s["mean"] * 0 + 0.25, s["var"] * 0 + 4.0, 10.0)
    y = pipe.normalize_obs(x).cpu().numpy()
    np.testing.assert_allclose(y, np.clip((x.cpu().numpy() - 0.25) / np.sqrt(4.0 + 1e-8), -10, 10), atol=1e-6)
    pipe.close(); sim.close(); ref.close()


@pytest.mark.gpu
def test_frame_stack_without_normalisation_is_exact():
    import torch

    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.sim import HlynrSim

    n = 515
    cfg = config.baseline_config("cfg2")
    cfg["max_steps"] = 25
    ref = HlynrSim(cfg, n_envs=n, seed=9, warn_dead=False)
    sim = HlynrSim(cfg, n_envs=n, seed=9, warn_dead=False)
    pipe = HlynrObsPipeline(sim, n_stack=4, norm_obs=False)
    stack = sb3_post.StackedObservations(n, 4, 26)
    np.testing.assert_array_equal(pipe.reset().cpu().numpy(), stack.reset(ref.reset().cpu().numpy()))
    rng = np.random.default_rng(2)
    for t in range(60):
        a = torch.as_tensor(rng.uniform(-1, 1, (n, 6)).astype(np.float32)).cuda()
        o, r, te, tr, tobs, _ = ref.step(a)
        dones = (te | tr).cpu().numpy().astype(bool)
        want, _ = stack.update(o.cpu().numpy(), dones, {})
        out = pipe.step(a)[0]
        np.testing.assert_array_equal(out.cpu().numpy(), want)
    pipe.close(); sim.close(); ref.close()
