"""Device-side rollout collection (include/hlynr_rollout.h, rollout.py) against the numpy restatement of
SB3's collect_rollouts / RolloutBuffer (oracle/sb3_post.py; parity unpinned: SB3 is not importable here)."""
import ctypes as C

import numpy as np
import pytest

from hlynr_intercept_b200 import config
from oracle import sb3_post


def test_gae_restatement_known_answer():
    """CPU: two steps, one env, hand-computed GAE."""
    r = np.array([[1.0], [2.0]], np.float32); v = np.array([[0.5], [0.25]], np.float32)
    starts = np.array([[1.0], [0.0]], np.float32)
    adv, ret = sb3_post.compute_returns_and_advantage(r, v, starts, np.array([4.0], np.float32), np.array([False]), 0.5, 0.5)
    d1 = 2.0 + 0.5 * 4.0 - 0.25
    d0 = 1.0 + 0.5 * 0.25 - 0.5
    np.testing.assert_allclose(adv[:, 0], [d0 + 0.25 * d1, d1], rtol=1e-6)
    np.testing.assert_allclose(ret, adv + v)


@pytest.mark.gpu
def test_gae_kernel_bit_exact_vs_sb3_restatement():
    import torch

    from hlynr_intercept_b200 import _lib

    L = _lib.load()
    rng = np.random.default_rng(0)
    T, N = 37, 1000
    r = rng.normal(0, 3, (T, N)).astype(np.float32); v = rng.normal(0, 2, (T, N)).astype(np.float32)
    starts = (rng.uniform(size=(T, N)) < 0.05).astype(np.float32)
    lv = rng.normal(0, 2, N).astype(np.float32); ld = rng.uniform(size=N) < 0.1
    want_adv, want_ret = sb3_post.compute_returns_and_advantage(r, v, starts, lv, ld, 0.99, 0.95)
    d = lambda x: torch.as_tensor(x).cuda()  # noqa: E731
    tr, tv, ts, tlv, tld = d(r), d(v), d(starts), d(lv), d(ld.astype(np.uint8))
    adv, ret = torch.empty_like(tr), torch.empty_like(tr)
    p = lambda x: C.c_void_p(x.data_ptr())  # noqa: E731
    _lib.check(L.hlynr_gae(p(tr), p(tv), p(ts), p(tlv), p(tld), T, N, 0.99, 0.95, p(adv), p(ret), 0, None))
    torch.cuda.synchronize()
    np.testing.assert_array_equal(adv.cpu().numpy(), want_adv)
    np.testing.assert_array_equal(ret.cpu().numpy(), want_ret)


@pytest.mark.gpu
def test_collector_matches_step_by_step_restatement():
    """One collect() of the device collector == the same loop written with the tensor API + numpy SB3 pieces."""
    import torch

    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.rollout import DeviceRolloutCollector, GaussianMlpPolicy
    from hlynr_intercept_b200.sim import HlynrSim

    n, T, k, gamma, lam = 2000, 48, 4, 0.99, 0.95
    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 20      # every env truncates at ticks 20, 40: the TimeLimit bootstrap path is exercised
    torch.manual_seed(0)
    pol = GaussianMlpPolicy(26 * k, net_arch=(64, 64), device="cuda")
    sims = [HlynrSim(cfg, n_envs=n, seed=3, warn_dead=False) for _ in range(2)]
    pipes = [HlynrObsPipeline(s, n_stack=k, training=True) for s in sims]
    col = DeviceRolloutCollector(pipes[0], pol, T, gamma=gamma, gae_lambda=lam, bootstrap_rows=n)
    torch.manual_seed(123)
    col.collect()
    torch.cuda.synchronize()
    assert int(col.overflow.item()) == 0
    # the same rollout, step by step
    torch.manual_seed(123)
    pipe = pipes[1]
    obs = pipe.reset().clone()
    starts = np.ones(n, np.float32)
    R, V, S = np.zeros((T, n), np.float32), np.zeros((T, n), np.float32), np.zeros((T, n), np.float32)
    n_boot = 0
    with torch.no_grad():
        for t in range(T):
            a, v, lp = pol(obs)
            np.testing.assert_array_equal(col.obs[t].cpu().numpy(), obs.cpu().numpy())
            np.testing.assert_array_equal(col.actions[t].cpu().numpy(), a.cpu().numpy())
            out, rew, te, tr, (records, counter, terminal) = pipe.step(a.clamp(-1, 1))
            rew = rew.cpu().numpy().copy()
            rec = pipe.done_records()
            term = terminal.cpu().numpy()
            for j, rr in enumerate(rec):
                if (rr["flags"] & 0x200) and not (rr["flags"] & 0x100):
                    tv = pol.value(torch.as_tensor(term[j:j + 1]).cuda()).cpu().numpy()[0]
                    rew[rr["env"]] += np.float32(gamma) * tv
                    n_boot += 1
            R[t], V[t], S[t] = rew, v.cpu().numpy(), starts
            starts = (te | tr).cpu().numpy().astype(np.float32)
            obs = out.clone()
        last_v = pol.value(obs).cpu().numpy()
    assert n_boot >= 2 * n
    np.testing.assert_allclose(col.rewards.cpu().numpy(), R, rtol=1e-6, atol=1e-6)
    np.testing.assert_array_equal(col.episode_starts.cpu().numpy(), S)
    np.testing.assert_allclose(col.values.cpu().numpy(), V, rtol=1e-5, atol=1e-6)
    adv, ret = sb3_post.compute_returns_and_advantage(col.rewards.cpu().numpy(), col.values.cpu().numpy(), S, last_v,
                                                      starts.astype(bool), gamma, lam)
    np.testing.assert_allclose(col.advantages.cpu().numpy(), adv, rtol=1e-5, atol=1e-4)
    np.testing.assert_allclose(col.returns.cpu().numpy(), ret, rtol=1e-5, atol=1e-4)
    for p_ in pipes:
        p_.close()
    for s in sims:
        s.close()


@pytest.mark.gpu
def test_graph_replay_equals_eager_collect():
    """collect() captured into a CUDA graph and replayed == the same collects issued eagerly (bit for bit)."""
    import torch

    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.rollout import DeviceRolloutCollector
    from hlynr_intercept_b200.sim import HlynrSim

    class Pol(torch.nn.Module):   # deterministic policy: no RNG, so eager and replayed runs are comparable
        def __init__(self):
            super().__init__()
            g = torch.Generator().manual_seed(0)
            self.w = torch.nn.Parameter(torch.randn(104, 6, generator=g).cuda() * 0.3)
            self.v = torch.nn.Parameter(torch.randn(104, generator=g).cuda() * 0.1)

        def value(self, obs):
            return obs @ self.v

        def forward(self, obs):
            a = torch.tanh(obs @ self.w)
            return a, obs @ self.v, -(a * a).sum(-1)

    n, k = 1500, 4
    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 30
    cols = []
    for _ in range(2):
        sim = HlynrSim(cfg, n_envs=n, seed=11, warn_dead=False)
        pipe = HlynrObsPipeline(sim, n_stack=k, training=True)
        cols.append(DeviceRolloutCollector(pipe, Pol(), n_steps=cols[0].graph_period() if cols else 12, bootstrap_rows=n))
    T = cols[0].graph_period()
    assert T % 12 == 0 and cols[0].T == T
    eager, graphed = cols
    graphed.capture()            # runs one eager collect (warm-up) before recording
    eager.collect()
    for rep in range(3):
        eager.collect()
        graphed.replay()
        torch.cuda.synchronize()
        for name in ("obs", "actions", "rewards", "episode_starts", "values", "advantages", "returns"):
            a, b = getattr(eager, name).cpu().numpy(), getattr(graphed, name).cpu().numpy()
            np.testing.assert_array_equal(a, b, err_msg=f"{name} after replay {rep}")
    se, sg = eager.sim.export_state(), graphed.sim.export_state()
    assert (se["steps"] == sg["steps"]).all() and (se["episode"] == sg["episode"]).all() and (se["ipos"] == sg["ipos"]).all()
    assert eager.sim.stats()["env_steps"] == graphed.sim.stats()["env_steps"] == 4 * T * n
    with pytest.raises(ValueError):
        DeviceRolloutCollector(graphed.pipe, Pol(), n_steps=T + 1).capture()


@pytest.mark.gpu
def test_episode_recorder_writes_the_reference_logger_format(tmp_path):
    import json

    import torch

    from hlynr_intercept_b200.episode_log import EpisodeRecorder
    from hlynr_intercept_b200.sim import HlynrSim

    cfg = config.baseline_config("cfg4")
    cfg["max_steps"] = 60
    sim = HlynrSim(cfg, n_envs=64, seed=2, warn_dead=False)
    sim.reset()
    rec = EpisodeRecorder(sim, str(tmp_path), env_index=5, metadata={"scenario": "medium"})
    rng = np.random.default_rng(0)
    done = False
    for t in range(80):
        a = rng.uniform(-1, 1, (64, 6)).astype(np.float32)
        obs, rew, te, tr, _, info = sim.step(torch.as_tensor(a).cuda(), auto_reset=False, want_info=True)
        done = rec.record_tick(a[5], float(rew[5]), bool(te[5]), bool(tr[5]), float(info["distance"][5]), bool(info["flags"][5] & 1))
        if done:
            break
    assert done
    lines = [json.loads(x) for x in open(tmp_path / "episodes" / "ep_000001.jsonl")]
    assert lines[0]["type"] == "header" and lines[0]["metadata"] == {"scenario": "medium"}
    states = [x for x in lines if x["type"] == "state"]
    assert len(states) == 2 * 60 and {x["entity_id"] for x in states} == {"interceptor", "missile"}
    assert len(states[0]["state"]["position"]) == 3 and len(states[0]["state"]["action"]) == 6
    foot = lines[-1]
    assert foot["type"] == "footer" and foot["outcome"] in ("intercepted", "failed") and foot["metrics"]["steps"] == 60
    sim.close()
