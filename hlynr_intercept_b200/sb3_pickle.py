"""vec_normalize.pkl without stable-baselines3 installed.

The reference saves `VecNormalize.save(path)` (rl_system/scripts/train_flat_ppo.py:477,530) and reads it back with
`VecNormalize.load(path, venv)` (rl_system/inference.py:163-170,454-468).  SB3's save is `pickle.dump(self, f)` of the wrapper
object (its __getstate__ drops `venv`, `class_attributes` and `returns`), so the file is a pickle that references

    stable_baselines3.common.vec_env.vec_normalize.VecNormalize      (instance, state = its __dict__)
    stable_baselines3.common.running_mean_std.RunningMeanStd         (obs_rms, ret_rms: mean, var, count)
    gymnasium.spaces.box.Box                                         (observation_space, action_space)

by module path.  `dump()` writes such a file from the device-side statistics of HlynrObsPipeline; `load()` reads one.  When SB3 /
gymnasium are importable the real classes are used; otherwise stub classes registered under the same module paths stand in for
the duration of the call, so the pickle byte stream names exactly the classes SB3 would.

PARITY UNPINNED: stable_baselines3 (rl_system/requirements.txt, unpinned 2.x) is not installable in the build image, so the
attribute set below restates SB3 2.x / gymnasium 0.29-1.x from their published sources and has not been loaded by a real
`VecNormalize.load` here.
"""
import contextlib
import pickle
import sys
import types

import numpy as np

_SB3_VN = "stable_baselines3.common.vec_env.vec_normalize"
_SB3_RMS = "stable_baselines3.common.running_mean_std"
_GYM_BOX = "gymnasium.spaces.box"


def _have_real():
    try:
        import gymnasium.spaces.box  # noqa: F401
        import stable_baselines3.common.vec_env.vec_normalize  # noqa: F401

        return True
    except Exception:
        return False


@contextlib.contextmanager
def _class_tree():
    """Yields (VecNormalize, RunningMeanStd, Box): the real classes, or stubs importable under SB3's / gymnasium's module paths."""
    if _have_real():
        from gymnasium.spaces.box import Box
        from stable_baselines3.common.running_mean_std import RunningMeanStd
        from stable_baselines3.common.vec_env.vec_normalize import VecNormalize

        yield VecNormalize, RunningMeanStd, Box
        return
    created = []

    def module(name):
        parts = name.split(".")
        for k in range(1, len(parts) + 1):
            full = ".".join(parts[:k])
            if full not in sys.modules:
                sys.modules[full] = types.ModuleType(full)
                created.append(full)
                if k > 1:
                    setattr(sys.modules[".".join(parts[:k - 1])], parts[k - 1], sys.modules[full])
        return sys.modules[name]

    def stub(mod_name, cls_name):
        m = module(mod_name)
        cls = getattr(m, cls_name, None)
        if cls is None:
            cls = type(cls_name, (), {"__module__": mod_name})
            cls.__qualname__ = cls_name
            setattr(m, cls_name, cls)
        return cls

    try:
        yield stub(_SB3_VN, "VecNormalize"), stub(_SB3_RMS, "RunningMeanStd"), stub(_GYM_BOX, "Box")
    finally:
        for name in reversed(created):
            sys.modules.pop(name, None)


def _new(cls, **state):
    obj = cls.__new__(cls)
    obj.__dict__.update(state)
    return obj


def _box(Box, low, high, shape, dtype=np.float32):
    dtype = np.dtype(dtype)
    lo, hi = np.full(shape, low, dtype=dtype), np.full(shape, high, dtype=dtype)
    # gymnasium.spaces.Box.__init__ (0.29 / 1.x): the attributes Space.__setstate__ restores with __dict__.update
    return _new(Box, dtype=dtype, _shape=tuple(shape), low=lo, high=hi, low_repr=str(low), high_repr=str(high),
                bounded_below=np.isfinite(lo), bounded_above=np.isfinite(hi), _np_random=None)


def dump(path, *, mean, var, count, ret_mean=0.0, ret_var=1.0, ret_count=1e-4, obs_shape, obs_low=-2.0, obs_high=1.0,
         act_shape=(6,), num_envs=1, clip_obs=10.0, clip_reward=10.0, gamma=0.99, epsilon=1e-8, training=True, norm_obs=True,
         norm_reward=False):
    """Writes what `VecNormalize.save(path)` writes (SB3 2.x vec_normalize.py: __getstate__ = __dict__ minus venv,
    class_attributes, returns)."""
    with _class_tree() as (VecNormalize, RunningMeanStd, Box):
        mean, var = np.asarray(mean, np.float64).reshape(obs_shape), np.asarray(var, np.float64).reshape(obs_shape)
        vn = _new(VecNormalize,
                  num_envs=int(num_envs), observation_space=_box(Box, obs_low, obs_high, tuple(obs_shape)),
                  action_space=_box(Box, -1.0, 1.0, tuple(act_shape)), render_mode=None,
                  norm_obs_keys=None, obs_spaces=None,
                  obs_rms=_new(RunningMeanStd, mean=mean, var=var, count=float(count)),
                  ret_rms=_new(RunningMeanStd, mean=np.float64(ret_mean), var=np.float64(ret_var), count=float(ret_count)),
                  clip_obs=float(clip_obs), clip_reward=float(clip_reward), gamma=float(gamma), epsilon=float(epsilon),
                  training=bool(training), norm_obs=bool(norm_obs), norm_reward=bool(norm_reward),
                  old_obs=np.array([]), old_reward=np.array([]))
        with open(path, "wb") as f:
            pickle.dump(vn, f)


def load(path):
    """Reads a vec_normalize.pkl (written by SB3 itself or by dump()) into a plain dict."""
    with _class_tree():
        with open(path, "rb") as f:
            vn = pickle.load(f)
    d = vn.__dict__
    orms, rrms = d["obs_rms"], d["ret_rms"]
    if isinstance(orms, dict):
        raise NotImplementedError("Dict observation spaces are not on the reference's path")
    return dict(mean=np.asarray(orms.mean, np.float64), var=np.asarray(orms.var, np.float64), count=float(orms.count),
                ret_mean=float(np.asarray(rrms.mean)), ret_var=float(np.asarray(rrms.var)), ret_count=float(rrms.count),
                clip_obs=float(d["clip_obs"]), clip_reward=float(d["clip_reward"]), gamma=float(d["gamma"]), epsilon=float(d["epsilon"]),
                training=bool(d["training"]), norm_obs=bool(d["norm_obs"]), norm_reward=bool(d["norm_reward"]),
                obs_shape=tuple(getattr(d["observation_space"], "_shape", None) or d["observation_space"].shape))
