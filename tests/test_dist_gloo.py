"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: contiguous global-id sharding, shard-invariant
trajectories, and the per-rollout all-reduce of the episode-statistics block.  The shards are played by the C
oracle here (no GPU in the build container); the CUDA equivalent is tests/test_cuda_parity.py::
test_sharding_invariance_and_rollout_equivalence."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from hlynr_intercept_b200 import abi, config
from hlynr_intercept_b200.dist import allreduce_stats, shard_range, summarize

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_shard_range_partitions_exactly():
    for n, w in [(1 << 20, 8), (10, 3), (7, 8), (4096, 2), (1, 1)]:
        got = [shard_range(n, r, w) for r in range(w)]
        assert got[0][0] == 0
        for (f0, c0), (f1, _) in zip(got, got[1:]):
            assert f0 + c0 == f1
        assert got[-1][0] + got[-1][1] == n
        assert max(c for _, c in got) - min(c for _, c in got) <= 1
    with pytest.raises(ValueError):
        shard_range(10, 3, 3)


def _worker(rank, world, port, n_total, T, seed, out_dir):
    sys.path.insert(0, ROOT)
    from oracle import draws, oracle

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    P, cur = config.resolve_config(config.baseline_config("cfg4"), warn_dead=False)
    first, count = shard_range(n_total, rank, world)
    sim = oracle.OracleBatch(P, cur, count, seed=seed, env_id_offset=first)
    obs = sim.reset()
    ids = np.arange(first, first + count)
    steps = np.zeros(count, np.int64)
    episode = np.zeros(count, np.int64)
    for t in range(T):
        a = draws.random_actions(seed, ids, episode, steps + 1)
        a[:, 0:3] = np.array([0, 0, -1], np.float32)  # dive: every episode ends by ground impact within T ticks
        obs, r, te, tr, _, _ = sim.step(a)
        done = (te | tr).astype(bool)
        episode += done
        steps = np.where(done, 0, steps + 1)
    st = sim.stats()
    local = torch.tensor([st.get(k, 0.0) for k in abi.STATS_FIELDS], dtype=torch.float64)
    total = allreduce_stats(local.clone())
    np.save(os.path.join(out_dir, f"obs_{rank}.npy"), obs)
    np.save(os.path.join(out_dir, f"stats_{rank}.npy"), np.array([total[k] for k in abi.STATS_FIELDS]))
    np.save(os.path.join(out_dir, f"local_{rank}.npy"), local.numpy())
    dist.destroy_process_group()


def test_two_rank_sharding_and_stats_allreduce(tmp_path):
    from oracle import draws, oracle

    n_total, T, seed, world = 24, 400, 5, 2
    port = 29500 + (os.getpid() % 500)
    mp.spawn(_worker, args=(world, port, n_total, T, seed, str(tmp_path)), nprocs=world, join=True)
    # single-process run of the same global batch
    P, cur = config.resolve_config(config.baseline_config("cfg4"), warn_dead=False)
    sim = oracle.OracleBatch(P, cur, n_total, seed=seed)
    obs = sim.reset()
    ids = np.arange(n_total)
    steps = np.zeros(n_total, np.int64)
    episode = np.zeros(n_total, np.int64)
    for t in range(T):
        a = draws.random_actions(seed, ids, episode, steps + 1)
        a[:, 0:3] = np.array([0, 0, -1], np.float32)
        obs, r, te, tr, _, _ = sim.step(a)
        done = (te | tr).astype(bool)
        episode += done
        steps = np.where(done, 0, steps + 1)
    want = sim.stats()
    sharded_obs = np.concatenate([np.load(tmp_path / f"obs_{r}.npy") for r in range(world)])
    assert (sharded_obs == obs).all(), "trajectories must not depend on the sharding"
    s0, s1 = np.load(tmp_path / "stats_0.npy"), np.load(tmp_path / "stats_1.npy")
    assert (s0 == s1).all(), "all-reduce result must be identical on every rank"
    loc = np.load(tmp_path / "local_0.npy") + np.load(tmp_path / "local_1.npy")
    np.testing.assert_allclose(s0, loc, rtol=1e-12)
    got = dict(zip(abi.STATS_FIELDS, s0))
    assert got["episodes"] == want["episodes"] > 0
    assert got["interceptor_crash"] == want["interceptor_crash"] > 0
    np.testing.assert_allclose(got["return_sum"], want["return_sum"], rtol=1e-9)
    summ = summarize(got)
    assert 0.0 <= summ["success_rate"] <= 1.0 and summ["mean_length"] > 0
