"""Launcher that lets the reference's trainers run UNCHANGED on the GPU simulator:

    python -m hlynr_intercept_b200.run train.py --config config.yaml
    python -m hlynr_intercept_b200.run inference.py --mode offline --model ...

The reference imports `DummyVecEnv` / `SubprocVecEnv` by name (rl_system/scripts/train_flat_ppo.py:22,371;
scripts/train_hrl_pretrain.py:349-361; inference.py:413).  Before running the target script this module patches
`stable_baselines3.common.vec_env.{DummyVecEnv,SubprocVecEnv}` with a factory: it calls env_fns[0]() once, unwraps
to the innermost env; if that is the reference's InterceptEnvironment with a configuration the accelerated path
covers, it returns `HlynrVecEnv(env.config, n_envs=len(env_fns))`, otherwise it falls through to the real class.

stable_baselines3 / gymnasium are not installed in the build image, so this file can only be exercised where they
are; the VecEnv contract itself is tested in tests/test_vec_env.py.
"""
import os
import runpy
import sys


def _innermost(env):
    seen = 0
    while hasattr(env, "env") and seen < 16:
        env = env.env
        seen += 1
    return env


def make_factory(real_cls, device=0, seed=1234, precision="fp32"):
    from .vec_env import HlynrVecEnv

    def factory(env_fns, *args, **kwargs):
        probe = env_fns[0]()
        inner = _innermost(probe)
        if type(inner).__name__ == "InterceptEnvironment" and hasattr(inner, "config"):
            try:
                venv = HlynrVecEnv(dict(inner.config), n_envs=len(env_fns), device=device, seed=seed, precision=precision)
            except NotImplementedError as e:  # outside the accelerated path (e.g. volley_size > 8): use the reference env
                print(f"[hlynr_intercept_b200] falling back to {real_cls.__name__}: {e}", file=sys.stderr)
            else:
                if hasattr(probe, "close"):
                    probe.close()
                return venv
        rest = list(env_fns)
        first = [probe]
        rest[0] = lambda: first.pop() if first else env_fns[0]()   # do not build env 0 twice
        return real_cls(rest, *args, **kwargs)

    return factory


def patch_sb3(device=None, seed=None, precision=None):
    import stable_baselines3.common.vec_env as vec_env

    device = int(os.environ.get("HLYNR_DEVICE", "0")) if device is None else device
    seed = int(os.environ.get("HLYNR_SEED", "1234")) if seed is None else seed
    precision = os.environ.get("HLYNR_PRECISION", "fp32") if precision is None else precision
    for name in ("DummyVecEnv", "SubprocVecEnv"):
        real = getattr(vec_env, name)
        fac = make_factory(real, device=device, seed=seed, precision=precision)
        setattr(vec_env, name, fac)
        sub = sys.modules.get("stable_baselines3.common.vec_env." + ("dummy_vec_env" if name == "DummyVecEnv" else "subproc_vec_env"))
        if sub is not None:
            setattr(sub, name, fac)


def main(argv=None):
    argv = list(sys.argv[1:] if argv is None else argv)
    if not argv:
        sys.exit("usage: python -m hlynr_intercept_b200.run <script.py> [script args...]")
    patch_sb3()
    script = argv[0]
    sys.argv = argv
    sys.path.insert(0, os.path.dirname(os.path.abspath(script)))
    runpy.run_path(script, run_name="__main__")


if __name__ == "__main__":
    main()
