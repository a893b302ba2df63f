"""fp32 build: statistical equivalence of episode outcomes with the UNMODIFIED reference running on its own NumPy
generators (no draw injection) -- north_star: "the fp32 build must also show statistical equivalence of episode
outcomes across randomized seeds".  The reference samples are tests/golden/stat_*.npz (oracle/gen_stat_golden.py):
96 first episodes per (config, scripted policy).  The CUDA side plays 4096 first episodes with its Philox draws."""
import json
import os

import numpy as np
import pytest
from scipy import stats as sps

from common import GOLDEN_DIR, CudaBatch
from hlynr_intercept_b200 import abi, config
from oracle import ref_harness

pytestmark = pytest.mark.gpu
CAUSES = ["intercepted", "hit_target", "interceptor_crash", "fuel_out", "missile_ground", "worsening", "timeout"]


def load_stat(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False)
    d = {k: z[k] for k in z.files}
    d["meta"] = json.loads(str(d["meta"]))
    return d


def play_first_episodes(env_cfg, policy, n=4096, seed=321, max_ticks=2100, training_step_count=None):
    P, cur = config.resolve_config(env_cfg, warn_dead=False)
    if training_step_count is not None:
        cur.set_training_step_count(training_step_count)
    sim = CudaBatch(P, cur, n, seed=seed)
    obs = sim.reset()
    pol = {"random": lambda: ref_harness.policy_random(5), "pursuit": ref_harness.policy_pursuit,
           "los_pn": lambda: ref_harness.policy_los_pn(5), "mixed": lambda: ref_harness.policy_mixed(5)}[policy]()
    out = dict(length=np.zeros(n), ret=np.zeros(n), min_distance=np.zeros(n), final_distance=np.zeros(n),
               cause=np.full(n, -1))
    open_ = np.ones(n, bool)
    for t in range(max_ticks):
        obs, rew, te, tr, tobs, info = sim.step(pol(t, obs))
        done = (te | tr).astype(bool) & open_
        if done.any():
            fl = info["flags"]
            ipz, mpz = info["interceptor_pos"][:, 2], info["missile_pos"][:, 2]
            cause = np.where(fl & abi.INFO_INTERCEPTED, 0,
                    np.where(~te.astype(bool), 6,
                    np.where(fl & abi.INFO_HIT_TARGET, 1,
                    np.where(ipz < 0, 2,
                    np.where(info["fuel_remaining"] <= 0, 3,
                    np.where(mpz <= 0, 4, 5))))))
            for k, src in (("length", info["episode_length"]), ("ret", info["episode_return"]),
                           ("min_distance", info["min_distance"]), ("final_distance", info["distance"])):
                out[k][done] = src[done]
            out["cause"][done] = cause[done]
            open_ &= ~done
        if not open_.any():
            break
    assert not open_.any(), "every env must finish its first episode within max_steps"
    return out


@pytest.mark.parametrize("name", ["stat_cfg4_random", "stat_cfg4_pursuit", "stat_cfg2_pursuit", "stat_hrl_los_pn",
                                  "stat_hrl_rotinv_mixed", "stat_volley3_pursuit"])
def test_episode_outcome_distributions_match_reference(name):
    ref = load_stat(name)
    meta = ref["meta"]
    got = play_first_episodes(meta["env_cfg"], meta["policy"], training_step_count=meta.get("training_step_count"))
    n_ref = len(ref["length"])
    # termination-cause proportions: within 4 sigma of the binomial error of the 96-episode reference sample
    for c, cname in enumerate(CAUSES):
        p_ref = float((ref["cause"] == c).mean())
        p_got = float((got["cause"] == c).mean())
        sigma = np.sqrt(max(p_got * (1 - p_got), 1e-4) / n_ref)
        assert abs(p_ref - p_got) <= 4 * sigma + 0.01, f"{cname}: reference {p_ref:.3f} vs cuda {p_got:.3f}"
    # distributions: two-sample Kolmogorov-Smirnov
    for k in ("length", "ret", "min_distance", "final_distance"):
        res = sps.ks_2samp(ref[k], got[k])
        assert res.pvalue > 1e-3, f"{k}: KS statistic {res.statistic:.3f}, p = {res.pvalue:.2e} " \
                                   f"(reference mean {ref[k].mean():.2f}, cuda mean {got[k].mean():.2f})"
