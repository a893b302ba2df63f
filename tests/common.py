"""Shared helpers for the parity tests."""
import json
import os

import numpy as np

from hlynr_intercept_b200 import config

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
GOLDEN_CASES = sorted(f[:-4] for f in os.listdir(GOLDEN_DIR) if f.endswith(".npz") and not f.startswith("stat_"))

STATE_FLOAT_FIELDS = ["ipos", "ivel", "quat", "mpos", "mvel", "wind", "thrust", "fuel", "fuel_used", "prev_d",
                      "last_d", "min_d", "kf_x", "kf_P", "T0", "base_cd", "peak"]
LOOSE_OBS = [9, 10, 11, 13, 16]
STATE_INT_FIELDS = ["steps", "worsen_count", "crossed", "kf_init", "onboard_delay", "episode"]

# north_star tolerances: fp32 build rtol 1e-3, fp64 build rtol 1e-5 (float64 flag -> tolerances).  Observation channels are
# normalised to [-1, 1], so their tolerance is absolute.  The worst errors observed on B200 are in profiles/parity_report.json
# (`python tools/parity_report.py` after a GPU run); every bound below is <= 10x the worst of its class there:
#   fp32 build: well-conditioned channels 4.0e-4 (golden replays, 47 fixtures) vs obs_atol 1e-3; ill-conditioned 5.3e-3 raw
#               (obs[13] before its sensitivity scaling) vs tti_atol 5e-3 scaled; reward floor 2.1e-4 vs reward_atol 2e-3
#   fp64 build: 1.7e-6 / 1.2e-5 (goldens / mixed sweep, no widening: the Kalman float32 -> float64 switch is reproduced) vs 1e-5;
#               ill-conditioned 6.0e-5 vs 1e-4; rewards within rtol 1e-5 with floor 0
#   dropped (low decision margin AND an actual disagreement): 0 on all 56 golden replays, the 4100-env runs and the full-size
#   windows; 1 env in the 30 x 1030-env fp32 mixed sweep.
TOL = {
    False: dict(rtol_state=1e-3, obs_atol=1e-3, reward_rtol=1e-3, reward_atol=2e-3, margin_tol=1e-4, tti_atol=5e-3),
    True: dict(rtol_state=1e-5, obs_atol=1e-5, reward_rtol=1e-5, reward_atol=1e-5, margin_tol=1e-6, tti_atol=1e-4),
}


def load_golden(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def golden_setup(g):
    """(HlynrParams, Curriculum) of a golden case."""
    meta = g["meta"]
    P, cur = config.resolve_config(meta["env_cfg"], warn_dead=False)
    if meta["training_step_count"] is not None:
        cur.set_training_step_count(meta["training_step_count"])
    got = cur.as_dict()
    for k, v in meta["curriculum"].items():
        assert abs(got[k] - v) < 1e-12, (k, got[k], v)
    return P, cur


REPORT_PATH = os.environ.get("HLYNR_PARITY_REPORT") or os.path.join(os.path.dirname(GOLDEN_DIR), os.pardir, "gpurun_out",
                                                                    "parity_report.jsonl")


class ParityLog:
    """Worst observed errors of one parity test.  Every `-m gpu` parity test appends its record to the JSON-lines file
    $HLYNR_PARITY_REPORT (default gpurun_out/parity_report.jsonl); tools/parity_report.py folds the records of a GPU run into
    the committed profiles/parity_report.json, which is what the tolerances in this directory are justified by."""

    def __init__(self, test, rtol, **meta):
        self.test, self.rtol, self.meta = test, float(rtol), meta
        self.obs_abs = np.zeros(26)      # worst |cuda - ref| per observation channel
        self.obs_floor = np.zeros(26)    # worst |cuda - ref| - rtol * |ref|: the absolute floor an rtol test would need
        self.fields = {}                 # name -> [worst abs, worst abs - rtol*|ref|, worst rel]
        self.dropped = 0
        self.compared = 0

    def obs(self, got, ref):
        if got.size == 0:
            return
        d = np.abs(got.astype(np.float64) - ref.astype(np.float64))
        for ch in (9, 11):               # roll / yaw over pi: -1 and +1 are the same angle
            d[:, ch] = np.minimum(d[:, ch], np.abs(2.0 - d[:, ch]))
        self.obs_abs = np.maximum(self.obs_abs, d.max(axis=0))
        self.obs_floor = np.maximum(self.obs_floor, (d - self.rtol * np.abs(ref)).max(axis=0))
        self.compared += got.shape[0]

    def field(self, name, got, ref):
        got, ref = np.asarray(got, np.float64), np.asarray(ref, np.float64)
        if got.size == 0:
            return
        d = np.abs(got - ref)
        w = self.fields.setdefault(name, [0.0, 0.0, 0.0])
        w[0] = max(w[0], float(d.max()))
        w[1] = max(w[1], float((d - self.rtol * np.abs(ref)).max()))
        w[2] = max(w[2], float((d / np.maximum(np.abs(ref), 1e-30))[np.abs(ref) > 1e-3].max(initial=0.0)))

    def write(self, **extra):
        rec = {"test": self.test, "rtol": self.rtol, "dropped": int(self.dropped), "env_ticks_compared": int(self.compared),
               "obs_abs": [float(f"{x:.3e}") for x in self.obs_abs], "obs_floor": [float(f"{max(x, 0.0):.3e}") for x in self.obs_floor],
               "fields": {k: {"abs": float(f"{v[0]:.3e}"), "floor": float(f"{max(v[1], 0.0):.3e}"), "rel": float(f"{v[2]:.3e}")}
                          for k, v in self.fields.items()}}
        rec.update(self.meta)
        rec.update(extra)
        try:
            os.makedirs(os.path.dirname(os.path.abspath(REPORT_PATH)), exist_ok=True)
            with open(REPORT_PATH, "a") as f:
                f.write(json.dumps(rec) + "\n")
        except OSError:
            pass
        return rec


def replay_against_golden(sim, g, rtol_state, obs_atol, reward_rtol, reward_atol, margin_fn=None, margin_tol=0.0,
                          tti_atol=None, log=None):
    """Replays the golden action sequence through `sim` (oracle or CUDA wrapper with the RefBatch call surface)
    and compares every output.  Integer/boolean outputs must match exactly, except for envs whose decision
    margin (as reported by margin_fn) dropped below margin_tol at some earlier tick: those are dropped from
    the comparison from that tick on and counted.  Returns a dict of worst-case errors."""
    T, n = g["obs"].shape[:2]
    obs0 = sim.reset()
    np.testing.assert_allclose(obs0, g["obs0"], rtol=0, atol=obs_atol)
    alive = np.ones(n, bool)
    worst = dict(obs=0.0, reward=0.0, dropped=0)
    term_k = 0
    tti_atol = obs_atol if tti_atol is None else tti_atol
    for t in range(T):
        obs, rew, te, tr, tobs, info = sim.step(g["actions"][t])
        if margin_fn is not None:
            m = margin_fn()
            newly = alive & (m < margin_tol)
            # a low-margin env is only dropped if it actually disagrees somewhere
            bad = newly & ((te != g["terminated"][t]) | (tr != g["truncated"][t]) | (info["flags"] != g["flags"][t])
                           | (np.abs(obs - g["obs"][t]).max(axis=1) > obs_atol))
            alive &= ~bad
            worst["dropped"] += int(bad.sum())
        a = alive
        if log is not None:
            log.dropped = worst["dropped"]
            log.obs(obs[a], g["obs"][t][a])
            log.field("reward", rew[a], g["reward"][t][a])
            for k in ("distance", "min_distance", "fuel_remaining", "fuel_used", "episode_return", "interceptor_pos", "missile_pos"):
                log.field(k, info[k][a], g[k][t][a])
        assert (te[a] == g["terminated"][t][a]).all(), f"terminated mismatch at t={t}"
        assert (tr[a] == g["truncated"][t][a]).all(), f"truncated mismatch at t={t}"
        assert (info["flags"][a] == g["flags"][t][a]).all(), \
            f"info flag mismatch at t={t}: {info['flags'][a]} vs {g['flags'][t][a]}"
        assert (info["steps"][a] == g["steps"][t][a]).all(), f"steps mismatch at t={t}"
        assert (info["episode_length"][a] == g["episode_length"][t][a]).all()
        d = np.abs(obs[a] - g["obs"][t][a])
        if d.size:
            # ill-conditioned channels get their own tolerance: euler angles / off-axis cosine (functions of the
            # quaternion, whose sin/cos ulps accumulate) and obs[13] = 1 - (range/closing)/100
            # obs[13] = 1 - x with x = range / (100 * closing): d obs[13] / d closing = 100 x^2 / range, so a closing-speed
            # difference that is invisible in obs[3:6] (tolerance obs_atol * max_velocity) is amplified by x^2 once the
            # closing speed fades (x -> 2 just before the channel clips to -1): the tolerance follows that sensitivity
            x = 1.0 - g["obs"][t][a][:, 13]
            d[:, 13] /= np.maximum(1.0, 4.0 * x * x)
            dl = d[:, LOOSE_OBS].max()
            d[:, LOOSE_OBS] = 0
            assert d.max() <= obs_atol, f"obs mismatch at t={t}: {d.max()} at {np.unravel_index(d.argmax(), d.shape)}"
            assert dl <= tti_atol, f"ill-conditioned obs channel mismatch at t={t}: {dl}"
            worst["obs"] = max(worst["obs"], float(d.max()))
        rg = g["reward"][t][a]
        err = np.abs(rew[a] - rg)
        assert (err <= reward_atol + reward_rtol * np.abs(rg)).all(), \
            f"reward mismatch at t={t}: {rew[a]} vs {rg}"
        if err.size:
            worst["reward"] = max(worst["reward"], float((err / (np.abs(rg) + 1e-3)).max()))
        for k in ("distance", "min_distance", "fuel_remaining", "fuel_used", "episode_return"):
            np.testing.assert_allclose(info[k][a], g[k][t][a], rtol=max(rtol_state, reward_rtol), atol=reward_atol,
                                       err_msg=f"info[{k}] at t={t}")
        for k in ("interceptor_pos", "missile_pos"):
            np.testing.assert_allclose(info[k][a], g[k][t][a], rtol=rtol_state, atol=rtol_state * 100,
                                       err_msg=f"info[{k}] at t={t}")
        if "missiles_intercepted" in g:   # volley-mode fixtures (environment.py:844-848)
            assert (info["missiles_intercepted"][a] == g["missiles_intercepted"][t][a]).all(), f"missiles_intercepted at t={t}"
            assert (info["missiles_remaining"][a] == g["missiles_remaining"][t][a]).all(), f"missiles_remaining at t={t}"
            np.testing.assert_allclose(info["missile_min_distances"][a], g["missile_min_distances"][t][a],
                                       rtol=max(rtol_state, reward_rtol), atol=reward_atol, err_msg=f"missile_min_distances at t={t}")
        done = (g["terminated"][t] | g["truncated"][t]).astype(bool)
        for i in np.nonzero(done)[0]:
            assert tuple(g["terminal_idx"][term_k]) == (t, i)
            if alive[i]:
                np.testing.assert_allclose(tobs[i], g["terminal_obs"][term_k], rtol=0, atol=max(obs_atol, tti_atol))
            term_k += 1
    st = sim.export_state()
    volley = "final_vpos" in g and g["meta"]["env_cfg"].get("volley_mode", False)
    for k in STATE_INT_FIELDS + (["vactive", "vcur", "vcount"] if volley else []):
        assert (st[k][alive] == g["final_" + k][alive]).all(), k
    for k in STATE_FLOAT_FIELDS + (["vpos", "vvel", "vmin"] if volley else []):
        ref = g["final_" + k][alive]
        if log is not None:
            log.field("final_" + k, st[k][alive], ref)
        scale = np.abs(ref).max() + 1e-6 if ref.size else 1.0
        np.testing.assert_allclose(st[k][alive], ref, rtol=rtol_state, atol=rtol_state * scale,
                                   err_msg=f"final state {k}")
    return worst


class CudaBatch:
    """HlynrSim behind the RefBatch / OracleBatch call surface (numpy in, numpy out), calling through the C ABI
    with device tensors."""

    def __init__(self, params, curriculum, n_envs, seed=1234, env_id_offset=0, float64=False, device=0):
        import torch
        from hlynr_intercept_b200.sim import HlynrSim

        self.torch = torch
        self.sim = HlynrSim(params=params, curriculum=curriculum, n_envs=n_envs, device=device, seed=seed,
                            env_id_offset=env_id_offset, precision="fp64" if float64 else "fp32")
        self.n = n_envs

    def reset(self, mask=None):
        m = None if mask is None else self.torch.as_tensor(np.asarray(mask, np.uint8))
        return self.sim.reset(m).cpu().numpy().copy()

    def step(self, actions, auto_reset=True):
        a = self.torch.as_tensor(np.ascontiguousarray(actions, np.float32)).to(self.sim.device)
        self.sim._alloc_out()["terminal_obs"].fill_(float("nan"))
        obs, rew, te, tr, tobs, info = self.sim.step(a, auto_reset=auto_reset, want_info=True)
        self.torch.cuda.synchronize()
        return (obs.cpu().numpy().copy(), rew.cpu().numpy().astype(np.float64), te.cpu().numpy().copy(),
                tr.cpu().numpy().copy(), tobs.cpu().numpy().copy(), {k: v.cpu().numpy().copy() for k, v in info.items()})

    def export_state(self):
        return self.sim.export_state()

    def stats(self, zero_after=False):
        return self.sim.stats(zero_after)


class Lockstep:
    """Steps a primary simulator and the oracle on the same actions; exposes the oracle's decision margins."""

    def __init__(self, primary, shadow):
        self.primary, self.shadow = primary, shadow
        self.margin = None

    def reset(self):
        self.shadow.reset()
        return self.primary.reset()

    def step(self, actions):
        self.shadow_out = self.shadow.step(actions)
        self.margin = self.shadow.margin
        return self.primary.step(actions)

    def export_state(self):
        return self.primary.export_state()
