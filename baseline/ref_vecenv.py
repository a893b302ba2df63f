"""CPU baseline: the UNMODIFIED reference InterceptEnvironment under a multiprocessing.Pipe clone of SB3's SubprocVecEnv.

BASELINE.md section 3 / north_star: "the reference CPU VecEnv (SubprocVecEnv) timed on the box's own host cores in the same
run, with the core count stated".  stable_baselines3 and gymnasium are not installable offline, so
  * the worker protocol below restates SB3's `_worker` (step with auto-reset + terminal_observation, reset, close) over
    multiprocessing.Pipe -- what `SubprocVecEnv(env_fns)` does at rl_system/scripts/train_hrl_pretrain.py:349-358;
  * `gymnasium` is the 2-class stub of SURVEY Appendix B (the reference only needs gym.Env.reset and spaces.Box).
The four reference modules are NOT part of this repository: `stage_reference()` (called by __graft_entry__.build() in the build
container, where /root/reference exists) copies them into the git-ignored baseline/_ref/, which travels to the GPU box with the
snapshot like the built .so files do.  Nothing here is imported by the product package.
"""
import multiprocessing as mp
import os
import shutil
import sys
import time
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
REF_DIR = os.path.join(HERE, "_ref", "rl_system")
REF_FILES = ["environment.py", "core.py", "physics_models.py", "physics_randomizer.py"]


def stage_reference(reference_root="/root/reference"):
    """Copies the reference's env modules into baseline/_ref/ (git-ignored).  Returns True if they are in place."""
    src = os.path.join(reference_root, "rl_system")
    if os.path.isfile(os.path.join(src, "environment.py")):
        os.makedirs(REF_DIR, exist_ok=True)
        for f in REF_FILES:
            shutil.copyfile(os.path.join(src, f), os.path.join(REF_DIR, f))
    return available()


def available():
    return all(os.path.isfile(os.path.join(REF_DIR, f)) for f in REF_FILES)


def _install_shim():
    try:
        import gymnasium  # noqa: F401  (prefer the real package when the box has it)
    except Exception:
        g = types.ModuleType("gymnasium")
        s = types.ModuleType("gymnasium.spaces")

        class Env:
            def reset(self, seed=None, options=None):
                return None

        class Box:
            def __init__(self, low, high, shape=None, dtype=None):
                self.low, self.high, self.shape, self.dtype = low, high, shape, dtype

        g.Env, s.Box, g.spaces = Env, Box, s
        sys.modules["gymnasium"] = g
        sys.modules["gymnasium.spaces"] = s
    if REF_DIR not in sys.path:
        sys.path.insert(0, REF_DIR)


def make_env(env_cfg, seed):
    _install_shim()
    from environment import InterceptEnvironment   # the reference's own class, unmodified

    env = InterceptEnvironment(env_cfg)
    env._hlynr_seed = seed
    return env


def _worker(remote, env_cfg, seed, envs_per_worker):
    """SB3 subproc_vec_env._worker restated: `step` auto-resets a finished env and returns the new episode's observation with
    info['terminal_observation']; one worker may host several envs (SB3 hosts one; >1 only amortises the pipe round trip)."""
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    envs = [make_env(env_cfg, seed + k) for k in range(envs_per_worker)]
    try:
        while True:
            cmd, data = remote.recv()
            if cmd == "step":
                out = []
                for env, a in zip(envs, data):
                    obs, reward, terminated, truncated, info = env.step(a)
                    done = terminated or truncated
                    info["TimeLimit.truncated"] = truncated and not terminated
                    if done:
                        info["terminal_observation"] = obs
                        obs, _ = env.reset()
                    out.append((obs, reward, done, info))
                remote.send(out)
            elif cmd == "reset":
                remote.send([env.reset(seed=env._hlynr_seed)[0] for env in envs])
            elif cmd == "close":
                remote.close()
                break
    except (EOFError, KeyboardInterrupt):
        pass


class PipeSubprocVecEnv:
    """Minimal SubprocVecEnv: one process per worker, step_async / step_wait over pipes."""

    def __init__(self, env_cfg, n_procs, envs_per_worker=1, seed=0):
        ctx = mp.get_context("fork")
        self.n_procs, self.epw = n_procs, envs_per_worker
        self.num_envs = n_procs * envs_per_worker
        self.remotes, self.procs = [], []
        for w in range(n_procs):
            parent, child = ctx.Pipe()
            p = ctx.Process(target=_worker, args=(child, env_cfg, seed + w * envs_per_worker, envs_per_worker), daemon=True)
            p.start()
            child.close()
            self.remotes.append(parent)
            self.procs.append(p)

    def reset(self):
        for r in self.remotes:
            r.send(("reset", None))
        return np.stack([o for r in self.remotes for o in r.recv()])

    def step_async(self, actions):
        for w, r in enumerate(self.remotes):
            r.send(("step", actions[w * self.epw:(w + 1) * self.epw]))

    def step_wait(self):
        res = [x for r in self.remotes for x in r.recv()]
        obs, rews, dones, infos = zip(*res)
        return np.stack(obs), np.asarray(rews, np.float32), np.asarray(dones), list(infos)

    def step(self, actions):
        self.step_async(actions)
        return self.step_wait()

    def close(self):
        for r in self.remotes:
            try:
                r.send(("close", None))
            except Exception:
                pass
        for p in self.procs:
            p.join(timeout=5)
            if p.is_alive():
                p.terminate()


def measure(env_cfg, seconds=10.0, n_procs=None, envs_per_worker=1, warmup_steps=20):
    """env-steps/s of the reference under the Pipe-clone SubprocVecEnv with random actions U(-1,1)^6 float32, plus the serial
    single-process (DummyVecEnv-style) figure.  Returns None when baseline/_ref/ is not staged."""
    if not available():
        return None
    n_procs = n_procs or os.cpu_count() or 1
    rng = np.random.default_rng(0)
    # serial, one env in this process
    env = make_env(env_cfg, 0)
    env.reset(seed=0)
    acts = rng.uniform(-1, 1, (256, 6)).astype(np.float32)
    for k in range(warmup_steps):
        env.step(acts[k % 256])
    t0, k = time.perf_counter(), 0
    budget = max(2.0, seconds * 0.25)
    while time.perf_counter() - t0 < budget:
        _, _, te, tr, _ = env.step(acts[k % 256])
        if te or tr:
            env.reset()
        k += 1
    serial = k / (time.perf_counter() - t0)
    # parallel
    venv = PipeSubprocVecEnv(env_cfg, n_procs, envs_per_worker)
    venv.reset()
    n = venv.num_envs
    pool = [rng.uniform(-1, 1, (n, 6)).astype(np.float32) for _ in range(8)]
    for k in range(warmup_steps):
        venv.step(pool[k % 8])
    t0, steps = time.perf_counter(), 0
    while time.perf_counter() - t0 < seconds:
        venv.step(pool[steps % 8])
        steps += 1
    dt = time.perf_counter() - t0
    venv.close()
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": n_procs, "kind": "reference",
            "serial_one_process": serial,
            "sample": f"unmodified reference InterceptEnvironment (baseline/_ref), {n_procs} worker processes x {envs_per_worker} env(s), "
                      f"{steps} vector steps in {dt:.1f} s, Pipe clone of SB3 SubprocVecEnv (stable_baselines3 / gymnasium not installable "
                      f"offline: SB3 _worker protocol restated, 2-class gymnasium stub), random actions U(-1,1)^6 float32, auto-reset"}


if __name__ == "__main__":
    sys.path.insert(0, os.path.dirname(HERE))
    from hlynr_intercept_b200 import config

    stage_reference()
    print(measure(config.baseline_config(sys.argv[1] if len(sys.argv) > 1 else "cfg4"), seconds=float(sys.argv[2]) if len(sys.argv) > 2 else 5.0))
