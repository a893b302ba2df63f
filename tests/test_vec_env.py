"""HlynrVecEnv: the Stable-Baselines3 VecEnv contract the reference's trainers rely on."""
import numpy as np
import pytest

from common import golden_setup, load_golden
from hlynr_intercept_b200 import config

pytestmark = pytest.mark.gpu


def make(n, cfg="cfg4", **kw):
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    return HlynrVecEnv(config.baseline_config(cfg), n_envs=n, seed=1234, warn_dead=False, **kw)


def test_spaces_and_shapes():
    v = make(3)
    assert v.num_envs == 3
    assert v.observation_space.shape == (26,) and v.observation_space.dtype == np.float32
    assert float(v.observation_space.low.min()) == -2.0 and float(v.observation_space.high.max()) == 1.0
    assert v.action_space.shape == (6,) and float(v.action_space.low.min()) == -1.0
    obs = v.reset()
    assert obs.shape == (3, 26) and obs.dtype == np.float32
    obs, rew, dones, infos = v.step(np.zeros((3, 6)))          # float64 zeros as helpers/check_missile_trajectory.py:58
    assert obs.shape == (3, 26) and rew.shape == (3,) and rew.dtype == np.float32 and dones.dtype == bool
    assert isinstance(infos, list) and len(infos) == 3
    for k in ("distance", "intercepted", "missile_hit_target", "fuel_remaining", "fuel_used", "clamped", "missile_pos",
              "interceptor_pos", "steps", "radar_detected", "radar_quality", "min_distance", "crossed_threshold"):
        assert k in infos[0], k
    assert infos[0]["steps"] == 1 and infos[0]["missile_pos"].shape == (3,)
    v.close()


def test_matches_reference_golden_through_the_vecenv_api():
    """Same trajectory as the golden fixture when driven through reset/step_async/step_wait, including the SB3
    auto-reset semantics (terminal_observation, TimeLimit.truncated, Monitor-style episode info)."""
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    g = load_golden("cfg4_f32_pursuit_long")
    meta = g["meta"]
    v = HlynrVecEnv(meta["env_cfg"], n_envs=meta["n_envs"], seed=meta["seed"], warn_dead=False)
    obs = v.reset()
    np.testing.assert_allclose(obs, g["obs0"], atol=1e-3)
    k = 0
    for t in range(g["obs"].shape[0]):
        v.step_async(g["actions"][t])
        obs, rew, dones, infos = v.step_wait()
        want_done = (g["terminated"][t] | g["truncated"][t]).astype(bool)
        assert (dones == want_done).all(), t
        d = np.abs(obs - g["obs"][t]); d[:, [9, 10, 11, 13, 16]] = 0
        assert d.max() < 1e-3
        np.testing.assert_allclose(rew, g["reward"][t], rtol=1e-3, atol=2e-3)
        for i in np.nonzero(want_done)[0]:
            inf = infos[i]
            np.testing.assert_allclose(inf["terminal_observation"], g["terminal_obs"][k], atol=5e-3)
            assert inf["TimeLimit.truncated"] == bool(g["truncated"][t][i] and not g["terminated"][t][i])
            assert inf["episode"]["l"] == int(g["episode_length"][t][i])
            np.testing.assert_allclose(inf["episode"]["r"], g["episode_return"][t][i], rtol=1e-3)
            assert inf["intercepted"] == bool(g["flags"][t][i] & 1)
            k += 1
        for i in np.nonzero(~want_done)[0]:
            assert "terminal_observation" not in infos[i] and "episode" not in infos[i]
    assert k == len(g["terminal_idx"]) and k > 0
    v.close()


def test_env_method_and_get_attr_surface():
    cfg = config.baseline_config("cfg4")
    cfg["curriculum"] = dict(enabled=True, initial_radius=100.0, final_radius=5.0, curriculum_steps=2000000)
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    v = HlynrVecEnv(cfg, n_envs=4, warn_dead=False)
    v.reset()
    assert v.env_method("get_current_intercept_radius") == [100.0] * 4
    v.env_method("set_training_step_count", 1000000)                        # scripts/train_flat_ppo.py:173-177
    assert abs(v.get_attr("get_current_intercept_radius")[0]() - 52.5) < 1e-9   # scripts/train_flat_ppo.py:221
    og = v.get_attr("observation_generator")[0]
    assert og.radar_beam_width == 60.0 and og.onboard_detection_reliability == 1.0
    st = v.get_attr("interceptor_state")[0]
    assert st["position"].shape == (3,) and st["orientation"].shape == (4,) and 0 < st["fuel"] <= 100.0
    ms = v.get_attr("missile_state", indices=[1, 2])
    assert len(ms) == 2 and ms[0]["velocity"].shape == (3,)
    assert v.env_is_wrapped(type("Monitor", (), {})) == [True] * 4
    with pytest.raises(AttributeError):
        v.env_method("render_everything")
    v.env_method("seed", 7)
    v.close()


def test_seed_reproducibility_and_curriculum_effect():
    a, b = make(16), make(16)
    a.seed(5); b.seed(5)
    oa, ob = a.reset(), b.reset()
    assert (oa == ob).all()
    act = np.random.default_rng(0).uniform(-1, 1, (16, 6)).astype(np.float32)
    for _ in range(20):
        ra, rb = a.step(act), b.step(act)
        assert (ra[0] == rb[0]).all() and (ra[1] == rb[1]).all()
    b.seed(6)
    assert not (b.reset() == a.reset()).all()
    a.close(); b.close()


def test_lazy_infos_large_batch():
    v = make(8192, lazy_infos=True, copy_outputs=False)
    v.reset()
    obs, rew, dones, infos = v.step(np.zeros((8192, 6), np.float32))
    assert len(infos) == 8192 and infos[0] is infos[1] and not dones.any() and len(infos.records) == 0
    v.close()


def test_host_pipeline_is_bit_identical_to_the_device_api_and_reports_done_records():
    """hlynr_step_host pipelines the shard in chunks over three streams (H2D | kernel | D2H) and returns finished
    episodes as compact records: obs / reward / done flags must equal the single-launch device API bit for bit
    (chunking invariance), and every record must equal the [N]-sized info arrays and terminal observation rows."""
    import torch

    from hlynr_intercept_b200.sim import HlynrSim
    from hlynr_intercept_b200 import abi

    n = 40001  # not a multiple of 128: ragged last chunk
    ref = HlynrSim(config.baseline_config("cfg4"), n_envs=n, seed=77, warn_dead=False)
    v = make(n, lazy_infos=True, copy_outputs=False)
    v.seed(77)
    v.sim.set_option("host_chunks", 5)
    ref.reset(); v.reset()
    ref.rollout(850, None); v.sim.rollout(850, None)  # age the episodes with the in-kernel random policy (same streams)
    rng = np.random.default_rng(3)
    total_done = 0
    for t in range(120):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        o, r, te, tr, tobs, info = ref.step(torch.as_tensor(a).cuda(), want_info=True)
        obs, rew, dones, infos = v.step(a)
        assert (obs == o.cpu().numpy()).all() and (rew == r.cpu().numpy()).all(), t
        want_done = (te | tr).cpu().numpy().astype(bool)
        assert (dones == want_done).all(), t
        rec = np.sort(infos.records, order="env")
        idx = np.nonzero(want_done)[0]
        assert (rec["env"] == idx).all(), t
        total_done += len(idx)
        if len(idx) == 0:
            continue
        inf = {k: x.cpu().numpy() for k, x in info.items()}
        assert (rec["terminal_obs"] == tobs.cpu().numpy()[idx]).all()
        for name in ("distance", "min_distance", "fuel_remaining", "fuel_used", "steps", "episode_return"):
            assert (rec[name] == inf[name][idx]).all(), name
        assert ((rec["flags"] & 0xff) == inf["flags"][idx]).all()
        assert (((rec["flags"] & abi.DONE_TERMINATED) != 0) == te.cpu().numpy()[idx].astype(bool)).all()
        assert (rec["missile_pos"] == inf["missile_pos"][idx]).all() and (rec["interceptor_pos"] == inf["interceptor_pos"][idx]).all()
        i = int(idx[0])  # lazily materialised dict of a finished env; others alias the shared empty dict
        d = infos[i]
        assert d["steps"] == int(inf["steps"][i]) and d["episode"]["l"] == d["steps"] and "terminal_observation" in d
        assert d is infos[i] and len(infos) == n
        other = int(np.nonzero(~want_done)[0][0])
        assert infos[other] == {}
    assert total_done > 50
    ref.close(); v.close()


def test_step_host_with_unpinned_buffers_and_terminal_rows():
    """The plain C-ABI call with caller-owned (unpinned) numpy buffers: staging copies + terminal rows scattered from
    the done records."""
    import ctypes as C

    from hlynr_intercept_b200 import _lib

    n = 3000
    a_env, b_env = make(n), make(n)
    a_env.reset(); b_env.reset()
    a_env.sim.rollout(900, None); b_env.sim.rollout(900, None)
    rng = np.random.default_rng(5)
    obs, rew = np.zeros((n, 26), np.float32), np.zeros(n, np.float32)
    te, tr = np.zeros(n, np.uint8), np.zeros(n, np.uint8)
    tobs = np.full((n, 26), np.nan, np.float32)
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    seen = 0
    for t in range(200):
        act = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        _lib.check(b_env.sim.L.hlynr_step_host(b_env.sim.h, p(act), p(obs), p(rew), p(te), p(tr), p(tobs), 1))
        o2, r2, d2, infos = a_env.step(act)
        np.testing.assert_array_equal(o2, obs, err_msg=f"tick {t}")
        np.testing.assert_array_equal(r2, rew)
        np.testing.assert_array_equal(d2, (te | tr) != 0)
        for i in np.nonzero(d2)[0]:
            assert (infos[i]["terminal_observation"] == tobs[i]).all()
            seen += 1
    assert seen > 5 and np.isnan(tobs).any()
    a_env.close(); b_env.close()


def test_volley_mode_info_fields():
    """inference.py --volley reads info['missiles_intercepted'], ['volley_size'], ['missiles_remaining'] (inference.py:561-585)."""
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    cfg = config.baseline_config("cfg4")
    cfg.update(volley_mode=True, volley_size=3)
    for lazy in (False, True):
        v = HlynrVecEnv(cfg, n_envs=6, seed=3, warn_dead=False, lazy_infos=lazy)
        v.reset()
        obs, rew, dones, infos = v.step(np.zeros((6, 6), np.float32))
        if not lazy:
            d = infos[0]
            assert d["volley_mode"] is True and d["volley_size"] == 3 and d["missiles_intercepted"] == 0
            assert d["missiles_remaining"] == 3 and len(d["missile_min_distances"]) == 3 and min(d["missile_min_distances"]) > 100.0
        v.sim.rollout(2100, None)       # every episode ends (max_steps 2000): done records carry the volley fields
        seen = 0
        for _ in range(3):
            obs, rew, dones, infos = v.step(np.zeros((6, 6), np.float32))
            for i in np.nonzero(dones)[0]:
                d = infos[int(i)]
                assert d["volley_size"] == 3 and 0 <= d["missiles_intercepted"] <= 3 and len(d["missile_min_distances"]) == 3
                assert "terminal_observation" in d and "episode" in d
                seen += 1
        v.close()


def test_outputs_of_a_step_survive_the_next_step():
    """Zero-copy mode returns views of page-locked buffers.  SB3's collect_rollouts reads the obs / dones / infos of step k
    after step k+1 has returned, so two output sets alternate: what step k returned must be untouched by step k+1 and may
    only be reused by step k+2 (pinned caller buffers, torch or library-owned, are DMA targets without a staging copy)."""
    n = 70000
    v = make(n, lazy_infos=True, copy_outputs=False)
    v.reset()
    v.sim.rollout(900, None)
    rng = np.random.default_rng(5)
    prev = None
    for t in range(40):
        obs, rew, dones, infos = v.step(rng.uniform(-1, 1, (n, 6)).astype(np.float32))
        assert dones.dtype == np.bool_ and (dones == (v._term | v._trunc).astype(bool)).all()
        if prev is not None:
            (o_view, r_view, d_view, i_view), (o_copy, r_copy, d_copy, rec_copy) = prev
            assert (o_view == o_copy).all() and (r_view == r_copy).all() and (d_view == d_copy).all(), t
            assert (i_view.records == rec_copy).all(), t
            assert not np.shares_memory(o_view, obs)
        prev = ((obs, rew, dones, infos), (obs.copy(), rew.copy(), dones.copy(), infos.records.copy()))
    v.close()


def test_step_host_with_pageable_and_page_locked_caller_buffers_agree():
    """hlynr_step_host through the raw C ABI, as INTEGRATION.md binds it: ordinary (pageable) numpy arrays are staged through
    the handle's pinned buffers, page-locked caller arrays (torch pin_memory) are DMA targets in place; both must return the
    same bits as the device API.  hlynr_host_done_buffer refuses pageable memory (the kernel stores into that buffer)."""
    import ctypes as C

    import torch

    from hlynr_intercept_b200 import _lib
    from hlynr_intercept_b200.sim import HlynrSim

    n = 70001
    sims = [HlynrSim(config.baseline_config("cfg4"), n_envs=n, seed=31, warn_dead=False) for _ in range(3)]
    for s in sims:
        s.reset()
        s.rollout(700, None)
    L = sims[0].L
    p = lambda x: x.ctypes.data_as(C.c_void_p)  # noqa: E731
    page = [np.zeros((n, 26), np.float32), np.zeros(n, np.float32), np.zeros(n, np.uint8), np.zeros(n, np.uint8)]
    pin_t = [torch.zeros((n, 26), dtype=torch.float32).pin_memory(), torch.zeros(n, dtype=torch.float32).pin_memory(),
             torch.zeros(n, dtype=torch.uint8).pin_memory(), torch.zeros(n, dtype=torch.uint8).pin_memory(),
             torch.zeros(n, dtype=torch.uint8).pin_memory()]
    pin = [t.numpy() for t in pin_t]
    with pytest.raises(_lib.HlynrError):
        _lib.check(L.hlynr_host_done_buffer(sims[2].h, p(np.zeros(n, np.uint8))))
    _lib.check(L.hlynr_host_done_buffer(sims[2].h, p(pin[4])))
    rng = np.random.default_rng(9)
    for t in range(60):
        a = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        o, r, te, tr, _, _ = sims[0].step(torch.as_tensor(a).cuda())
        _lib.check(L.hlynr_step_host(sims[1].h, p(a), p(page[0]), p(page[1]), p(page[2]), p(page[3]), None, 1))
        _lib.check(L.hlynr_step_host(sims[2].h, p(a), p(pin[0]), p(pin[1]), p(pin[2]), p(pin[3]), None, 1))
        want = (o.cpu().numpy(), r.cpu().numpy(), te.cpu().numpy(), tr.cpu().numpy())
        for k in range(4):
            assert (page[k] == want[k]).all() and (pin[k] == want[k]).all(), (t, k)
        assert (pin[4] == (want[2] | want[3])).all(), t
    for s in sims:
        s.close()


def test_obs_dim_17_is_the_leading_radar_channels():
    """HlynrVecEnv(obs_dim=17): the 17-D radar layout (rl_system/hrl/observation_schema.py:13-46) = obs[0:17] of the 26-D vector,
    bit for bit, through the host path (pipelined chunks, ragged last tile) and the tensor API; terminal observations too."""
    import torch
    from hlynr_intercept_b200.sim import HlynrSim

    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 30
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    n = 70001   # > 65536: zero-copy outputs, several chunks
    a26 = HlynrVecEnv(cfg, n_envs=n, seed=5, warn_dead=False)
    a17 = HlynrVecEnv(cfg, n_envs=n, seed=5, warn_dead=False, obs_dim=17)
    assert a17.observation_space.shape == (17,)
    o26, o17 = a26.reset(), a17.reset()
    assert o17.shape == (n, 17) and (o17 == o26[:, :17]).all()
    rng = np.random.default_rng(0)
    finished = 0
    for t in range(40):
        act = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        o26, r26, d26, i26 = a26.step(act)
        o17, r17, d17, i17 = a17.step(act)
        assert (o17 == o26[:, :17]).all() and (r17 == r26).all() and (d17 == d26).all(), t
        for i in np.nonzero(d26)[0][:5]:
            assert (i17[i]["terminal_observation"] == i26[i]["terminal_observation"][:17]).all()
            assert i17[i]["terminal_observation"].shape == (17,)
            finished += 1
    assert finished > 0
    a26.close(); a17.close()
    s26 = HlynrSim(cfg, n_envs=333, seed=5, warn_dead=False)
    s17 = HlynrSim(cfg, n_envs=333, seed=5, warn_dead=False, obs_dim=17)
    assert (s17.reset().cpu() == s26.reset().cpu()[:, :17]).all()
    for t in range(35):
        act = torch.rand(333, 6, device="cuda") * 2 - 1
        x26, x17 = s26.step(act), s17.step(act)
        assert x17[0].shape == (333, 17) and (x17[0].cpu() == x26[0].cpu()[:, :17]).all()
        done = (x26[2] | x26[3]).bool().cpu()
        if done.any():
            assert (x17[4].cpu()[done] == x26[4].cpu()[done][:, :17]).all()
    s26.close(); s17.close()


def test_radar_quality_is_zero_while_the_onboard_delay_buffer_fills():
    """info['radar_quality'] (environment.py:840) is last_detection_info['radar_quality']: 0.0 during the first onboard_delay
    steps of every episode ('sensor_delay_initialization', core.py:579-584), the configured quality afterwards."""
    cfg = config.baseline_config("cfg4")   # sensor delay 30 ms = 3 samples
    cfg["max_steps"] = 7
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    v = HlynrVecEnv(cfg, n_envs=4, seed=3, warn_dead=False)
    q = float(v.sim.params.radar_quality)
    delay = int(v.sim.params.onboard_delay)
    assert delay == 3 and q > 0
    v.reset()
    for t in range(1, 20):
        _, _, dones, infos = v.step(np.zeros((4, 6), np.float32))
        steps = infos[0]["steps"]
        assert steps == (t - 1) % 7 + 1
        assert infos[0]["radar_quality"] == (0.0 if steps < delay else q), (t, steps, infos[0]["radar_quality"])
    v.close()
    # without sensor delays the quality is there from the first step
    v = HlynrVecEnv(config.baseline_config("cfg2"), n_envs=2, seed=3, warn_dead=False)
    v.reset()
    assert v.step(np.zeros((2, 6), np.float32))[3][0]["radar_quality"] == float(v.sim.params.radar_quality)
    v.close()


def test_lazy_infos_expose_the_per_env_info_arrays_on_request():
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    cfg = config.baseline_config("cfg4")
    n = 5000
    lazy = HlynrVecEnv(cfg, n_envs=n, seed=8, warn_dead=False, lazy_infos=True, info_arrays=True)
    eager = HlynrVecEnv(cfg, n_envs=n, seed=8, warn_dead=False, lazy_infos=False)
    plain = HlynrVecEnv(cfg, n_envs=n, seed=8, warn_dead=False, lazy_infos=True)
    lazy.reset(); eager.reset(); plain.reset()
    rng = np.random.default_rng(1)
    for t in range(3):
        act = rng.uniform(-1, 1, (n, 6)).astype(np.float32)
        _, _, _, li = lazy.step(act)
        _, _, _, ei = eager.step(act)
        _, _, _, pi = plain.step(act)
    arr = li.arrays
    assert arr["distance"].shape == (n,) and arr["interceptor_pos"].shape == (n, 3)
    for i in (0, 17, n - 1):
        assert li[i].keys() == ei[i].keys()
        for k in ("distance", "fuel_remaining", "steps", "radar_detected", "radar_quality", "min_distance"):
            assert li[i][k] == ei[i][k], (i, k)
        assert (li[i]["missile_pos"] == ei[i]["missile_pos"]).all()
        assert arr["distance"][i] == ei[i]["distance"]
    assert pi[0] == {}   # not requested: unfinished envs share one empty dict
    with pytest.raises(Exception):
        pi.arrays
    lazy.close(); eager.close(); plain.close()


def test_seed_takes_effect_at_the_next_reset():
    """SB3 VecEnv.seed semantics: stored, applied by the next reset(); running episodes keep their draws."""
    a, b = make(64), make(64)
    a.reset(); b.reset()
    rng = np.random.default_rng(2)
    acts = rng.uniform(-1, 1, (12, 64, 6)).astype(np.float32)
    for t in range(6):
        if t == 3:
            a.seed(777)
        oa, ob = a.step(acts[t])[0], b.step(acts[t])[0]
        assert (oa == ob).all(), t          # mid-episode: unchanged
    ra, rb = a.reset(), b.reset()
    assert not (ra == rb).all()              # the new seed keys the episodes that start at the reset
    c = make(64)   # same draws as `a` after its re-seeded reset: seed 777, second episode of every env
    c.reset()
    c.seed(777)
    assert (c.reset() == ra).all()
    a.close(); b.close(); c.close()


def test_step_graph_replays_bit_identically():
    """HlynrSim.capture_steps: ring_period() ticks in ONE CUDA graph == the same ticks launched one by one; a replay at the
    wrong ring phase is refused."""
    import torch
    from hlynr_intercept_b200 import _lib
    from hlynr_intercept_b200.sim import HlynrSim

    for name in ("cfg2", "cfg4"):
        cfg = config.baseline_config(name)
        n = 4096
        a = HlynrSim(cfg, n_envs=n, seed=11, warn_dead=False)
        b = HlynrSim(cfg, n_envs=n, seed=11, warn_dead=False)
        a.reset(); b.reset()
        a.rollout(1000, None, want_obs=False); b.rollout(1000, None, want_obs=False)
        T = a.ring_period()
        assert T in (6, 12)
        while a.tick_count() % T:
            act = torch.zeros(n, 6, device="cuda")
            a.step(act); b.step(act)
        g = torch.Generator(device="cuda"); g.manual_seed(3)
        bufs = [torch.empty(n, 6, device="cuda") for _ in range(T)]
        sg = a.capture_steps(bufs)
        for rep in range(3):
            for k in range(T):
                bufs[k].copy_(torch.rand(n, 6, device="cuda", generator=g) * 2 - 1)
            out = sg.replay()
            for k in range(T):
                ref = b.step(bufs[k])
            torch.cuda.synchronize()
            for x, y in zip(out[:4], ref[:4]):
                assert (x == y).all()
        sa, sb = a.export_state(), b.export_state()
        for k in sa:
            assert (sa[k] == sb[k]).all(), k
        assert a.stats() == b.stats()
        a.step(bufs[0])
        with pytest.raises(_lib.HlynrError):
            sg.replay()
        a.close(); b.close()


@pytest.mark.gpu
def test_every_documented_option_is_accepted():
    """hlynr_set_option knows every option name the headers, tools and docs use (a clean-up once dropped two of them silently)."""
    from hlynr_intercept_b200 import config
    from hlynr_intercept_b200.sim import HlynrSim

    sim = HlynrSim(config.baseline_config("cfg4"), n_envs=256, warn_dead=False)
    for name, value in (("specialise", 1), ("host_info", 1), ("prefetch_waves", 1), ("obs_dim", 26), ("host_chunks", 0), ("pdl", 1),
                        ("host_chunk_growth", 12), ("host_threads", 0)):
        sim.set_option(name, value)
    with pytest.raises(Exception):
        sim.set_option("no_such_option", 1)
    sim.close()
