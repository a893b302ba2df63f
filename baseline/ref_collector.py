"""CPU counterpart of BASELINE configuration 5 ("end-to-end train.py PPO rollout collection ... compared with SubprocVecEnv on
host cores"): the loop SB3's `collect_rollouts` runs for the reference trainer (rl_system/scripts/train_flat_ppo.py:353-448),
on the host:

    SubprocVecEnv(unmodified InterceptEnvironment x n_procs)   baseline/ref_vecenv.py (Pipe clone; SB3 is not installable offline)
    -> VecFrameStack(4) -> VecNormalize(norm_obs, clip 10)     oracle/sb3_post.py (numpy restatement of SB3 2.x, parity unpinned)
    -> policy forward on the CPU (torch, fp32): the reference network 104 -> 512 -> 512 -> 256 + LayerNorm + ReLU, heads
    -> TimeLimit bootstrap, rollout buffer, GAE(lambda) at the end

It is the thing `DeviceRolloutCollector` replaces; bench.py times both on the same box (`cfg5` block).  Test / measurement
infrastructure only: nothing in the product package imports it.
"""
import os
import time

import numpy as np


def collect(env_cfg, seconds=10.0, n_procs=None, envs_per_worker=1, gamma=0.99, gae_lambda=0.95, n_stack=4, torch_threads=None):
    """Collects transitions for ~`seconds` and returns the env-steps/s of the whole loop (policy + envs + wrappers + GAE)."""
    import torch

    from baseline import ref_vecenv
    from hlynr_intercept_b200.policy import ReferenceActorCritic
    from oracle import sb3_post

    if not ref_vecenv.available():
        return {"unavailable": "baseline/_ref/ is not staged (run __graft_entry__.build() where /root/reference exists)"}
    n_procs = n_procs or os.cpu_count() or 1
    if torch_threads:
        torch.set_num_threads(torch_threads)
    torch.manual_seed(0)
    net = ReferenceActorCritic(device="cpu")
    venv = ref_vecenv.PipeSubprocVecEnv(env_cfg, n_procs, envs_per_worker)
    n = venv.num_envs
    stack = sb3_post.StackedObservations(n, n_stack, 26)
    norm = sb3_post.VecNormalize(n, (26 * n_stack,), training=True, gamma=gamma)
    obs = norm.reset(stack.reset(venv.reset()).copy())
    starts = np.ones(n, np.float32)
    buf = {k: [] for k in ("obs", "actions", "rewards", "starts", "values", "logp")}
    t0 = time.perf_counter()
    t_policy = t_env = 0.0
    steps = 0
    with torch.no_grad():
        while time.perf_counter() - t0 < seconds or steps < 8:
            ta = time.perf_counter()
            a, v, lp = net(torch.from_numpy(obs))
            a, v, lp = a.numpy(), v.numpy(), lp.numpy()
            tb = time.perf_counter()
            new_obs, rew, dones, infos = venv.step(np.clip(a, -1.0, 1.0))
            info_map = {i: infos[i] for i in np.nonzero(dones)[0]}
            stacked, info_map = stack.update(new_obs, dones, info_map)
            nobs, nrew, _, info_map = norm.step(stacked.copy(), rew, dones, info_map)
            rew = np.asarray(nrew, np.float32).copy()
            for i, inf in info_map.items():   # collect_rollouts: rewards[i] += gamma * V(terminal_observation) on time-outs
                if inf.get("TimeLimit.truncated", False):
                    rew[i] += gamma * float(net.value(torch.from_numpy(inf["terminal_observation"][None].astype(np.float32)))[0])
            tc = time.perf_counter()
            for k, x in zip(buf, (obs, a, rew, starts, v, lp)):
                buf[k].append(x)
            obs, starts = nobs, dones.astype(np.float32)
            t_policy += tb - ta
            t_env += tc - tb
            steps += 1
        last_values = net.value(torch.from_numpy(obs)).numpy()
    adv, ret = sb3_post.compute_returns_and_advantage(np.stack(buf["rewards"]), np.stack(buf["values"]), np.stack(buf["starts"]),
                                                      last_values, starts.astype(bool), gamma, gae_lambda)
    dt = time.perf_counter() - t0
    venv.close()
    assert np.isfinite(adv).all() and np.isfinite(ret).all()
    return {"value": n * steps / dt, "unit": "env-steps/s", "cores": n_procs, "envs": n, "n_steps": steps, "seconds": dt,
            "policy_share": t_policy / dt, "env_and_wrappers_share": t_env / dt,
            "what": "SB3 collect_rollouts on the host: Pipe-clone SubprocVecEnv of the unmodified reference env (one process per core) -> "
                    "numpy VecFrameStack(4) + VecNormalize -> torch fp32 policy on the CPU (reference network) -> TimeLimit bootstrap -> "
                    "GAE (stable_baselines3 itself is not installable offline: its loop is restated, parity unpinned)"}


if __name__ == "__main__":
    import sys

    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    from hlynr_intercept_b200 import config

    print(collect(config.baseline_config("cfg4"), seconds=float(sys.argv[1]) if len(sys.argv) > 1 else 5.0))
