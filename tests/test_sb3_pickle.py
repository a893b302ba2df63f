"""vec_normalize.pkl round trip without stable_baselines3 (hlynr_intercept_b200/sb3_pickle.py): the byte stream names SB3's
and gymnasium's classes by their module paths, loads back to the same statistics, and leaves no stub modules behind.
PARITY UNPINNED: a real `VecNormalize.load` cannot be run in this image (SB3 is not installable offline)."""
import pickle
import pickletools
import sys

import numpy as np
import pytest

from hlynr_intercept_b200 import sb3_pickle


def test_round_trip_and_class_paths(tmp_path):
    rng = np.random.default_rng(0)
    before = set(sys.modules)
    mean, var = rng.normal(size=104), rng.uniform(0.1, 2.0, 104)
    p = tmp_path / "vec_normalize.pkl"
    sb3_pickle.dump(p, mean=mean, var=var, count=12345.0, ret_mean=0.5, ret_var=2.0, ret_count=99.0, obs_shape=(104,), num_envs=16,
                    clip_obs=10.0, gamma=0.99, epsilon=1e-8, training=False)
    d = sb3_pickle.load(p)
    assert (d["mean"] == mean).all() and (d["var"] == var).all() and d["count"] == 12345.0
    assert (d["ret_mean"], d["ret_var"], d["ret_count"]) == (0.5, 2.0, 99.0)
    assert d["obs_shape"] == (104,) and d["clip_obs"] == 10.0 and d["training"] is False and d["norm_obs"] is True
    names = {arg for op, arg, _ in pickletools.genops(p.read_bytes()) if op.name in ("GLOBAL", "STACK_GLOBAL", "SHORT_BINUNICODE", "BINUNICODE")
             if isinstance(arg, str)}
    for want in ("stable_baselines3.common.vec_env.vec_normalize", "VecNormalize", "stable_baselines3.common.running_mean_std",
                 "RunningMeanStd", "gymnasium.spaces.box", "Box"):
        assert want in names, want
    for k in ("venv", "class_attributes", "returns"):   # VecNormalize.__getstate__ drops these
        assert k not in names
    if not sb3_pickle._have_real():
        assert set(sys.modules) == before   # the stub modules are gone again
        with pytest.raises((ModuleNotFoundError, AttributeError)):
            pickle.loads(p.read_bytes())   # the file really needs SB3's classes (or the stub tree) to load


@pytest.mark.gpu
def test_pipeline_statistics_survive_the_pickle(tmp_path):
    import torch
    from hlynr_intercept_b200 import config
    from hlynr_intercept_b200.post import HlynrObsPipeline
    from hlynr_intercept_b200.sim import HlynrSim

    cfg = config.baseline_config("cfg4")
    sim = HlynrSim(cfg, n_envs=512, seed=3, warn_dead=False)
    pipe = HlynrObsPipeline(sim, n_stack=4, training=True)
    pipe.reset()
    for _ in range(20):
        pipe.step(torch.rand(512, 6, device="cuda") * 2 - 1)
    p = tmp_path / "vec_normalize.pkl"
    pipe.save_sb3_pickle(str(p))
    want = pipe.get_stats()
    sim2 = HlynrSim(cfg, n_envs=64, seed=4, warn_dead=False)
    pipe2 = HlynrObsPipeline(sim2, n_stack=4, training=False)   # inference.py:468: frozen statistics
    pipe2.load_sb3_pickle(str(p))
    got = pipe2.get_stats()
    for k in want:
        assert np.array_equal(np.asarray(want[k]), np.asarray(got[k])), k
    pipe.close(); pipe2.close(); sim.close(); sim2.close()
