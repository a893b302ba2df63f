"""hlynr_intercept_b200.run: the launcher that lets the reference's trainers run unchanged by patching
stable_baselines3.common.vec_env.{DummyVecEnv,SubprocVecEnv}.  stable_baselines3 is not installed in the build
image, so a stand-in package with the two class names is put into sys.modules."""
import sys
import textwrap
import types

import numpy as np
import pytest

from hlynr_intercept_b200 import config, run


class _RealVecEnv:
    def __init__(self, env_fns, *a, **k):
        self.envs = [f() for f in env_fns]
        self.num_envs = len(self.envs)


class InterceptEnvironment:   # same class name as rl_system/environment.py:17; the factory reads .config
    built = 0

    def __init__(self, cfg):
        type(self).built += 1
        self.config = cfg


class _Monitor:               # Monitor(InterceptEnvironment(cfg)) as in scripts/train_flat_ppo.py:344-350
    def __init__(self, env):
        self.env = env

    def close(self):
        pass


class _OtherEnv:
    built = 0

    def __init__(self):
        type(self).built += 1


@pytest.fixture
def fake_sb3(monkeypatch):
    pkg, common, vec = (types.ModuleType("stable_baselines3"), types.ModuleType("stable_baselines3.common"),
                        types.ModuleType("stable_baselines3.common.vec_env"))
    vec.DummyVecEnv = type("DummyVecEnv", (_RealVecEnv,), {})
    vec.SubprocVecEnv = type("SubprocVecEnv", (_RealVecEnv,), {})
    pkg.common, common.vec_env = common, vec
    for name, m in (("stable_baselines3", pkg), ("stable_baselines3.common", common), ("stable_baselines3.common.vec_env", vec)):
        monkeypatch.setitem(sys.modules, name, m)
    return vec


def test_other_environments_fall_through_to_the_real_class(fake_sb3):
    real = fake_sb3.DummyVecEnv
    run.patch_sb3(device=0, seed=1, precision="fp32")
    assert fake_sb3.DummyVecEnv is not real and fake_sb3.SubprocVecEnv is not None
    _OtherEnv.built = 0
    v = fake_sb3.DummyVecEnv([_OtherEnv] * 5)
    assert isinstance(v, real) and v.num_envs == 5
    assert _OtherEnv.built == 5   # the probe of env 0 is reused, not built twice
    assert isinstance(fake_sb3.DummyVecEnv, type) and issubclass(fake_sb3.DummyVecEnv, real)   # still a class: isinstance / subclassing work
    assert fake_sb3.DummyVecEnv.__name__ == "DummyVecEnv"


def test_env_fns_with_different_configs_fall_back(fake_sb3, capsys):
    """The GPU batch shares ONE configuration: env_fns whose configs differ must go to the reference's own vec env."""
    real = fake_sb3.SubprocVecEnv
    run.patch_sb3(device=0, seed=1, precision="fp32")
    a, b = config.baseline_config("cfg4"), config.baseline_config("cfg2")
    InterceptEnvironment.built = 0
    fns = [lambda: _Monitor(InterceptEnvironment(a))] * 3 + [lambda: _Monitor(InterceptEnvironment(b))]
    v = fake_sb3.SubprocVecEnv(fns)
    assert isinstance(v, real) and v.num_envs == 4
    assert InterceptEnvironment.built == 4    # probes reused
    assert "different configuration" in capsys.readouterr().err


@pytest.mark.gpu
def test_reference_environments_become_the_gpu_vecenv(fake_sb3, tmp_path):
    from hlynr_intercept_b200.vec_env import HlynrVecEnv

    run.patch_sb3(device=0, seed=7, precision="fp32")
    cfg = config.baseline_config("cfg4")
    InterceptEnvironment.built = 0
    for name in ("DummyVecEnv", "SubprocVecEnv"):
        v = getattr(fake_sb3, name)([lambda: _Monitor(InterceptEnvironment(cfg))] * 24)
        assert isinstance(v, HlynrVecEnv) and v.num_envs == 24 and v.config == cfg
        obs = v.reset()
        obs, rew, dones, infos = v.step(np.zeros((24, 6), np.float32))
        assert obs.shape == (24, 26) and rew.shape == (24,) and len(infos) == 24
        assert v.env_method("get_current_intercept_radius")[0] > 0
        v.close()
    assert InterceptEnvironment.built == 2 * 8   # a few probes (first, last, 6 in between) per vec env, never one reference env per GPU env
    # the way it is used: python -m hlynr_intercept_b200.run <unchanged script> ...
    script = tmp_path / "train_like.py"
    script.write_text(textwrap.dedent("""
        import sys
        import numpy as np
        from stable_baselines3.common.vec_env import DummyVecEnv
        from test_launcher import InterceptEnvironment, _Monitor
        from hlynr_intercept_b200 import config
        cfg = config.baseline_config("cfg2")
        venv = DummyVecEnv([lambda: _Monitor(InterceptEnvironment(cfg)) for _ in range(int(sys.argv[1]))])
        venv.reset()
        for _ in range(5):
            obs, rew, dones, infos = venv.step(np.zeros((venv.num_envs, 6), np.float32))
        open(sys.argv[2], "w").write(f"{type(venv).__name__} {venv.num_envs} {obs.shape[1]}")
    """))
    out = tmp_path / "out.txt"
    argv = list(sys.argv)
    try:
        run.main([str(script), "12", str(out)])
    finally:
        sys.argv = argv
    assert out.read_text() == "HlynrVecEnv 12 26"
