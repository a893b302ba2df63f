"""Multi-GPU plumbing: one process per GPU (torch.distributed), environments sharded by GLOBAL id.

Environments never interact (each InterceptEnvironment owns all its state; only the curriculum scalars are
global, rl_system/environment.py:269-272), so the batch is split into contiguous global-id ranges with NO
per-step communication.  The Philox key uses the global env id, hence trajectories are bit-identical for any
GPU count.  The only collective is one all-reduce (sum) of the 16-double episode-statistics block per rollout
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

from . import abi


def env_rank_world():
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(n_total, rank, world):
    """Contiguous global-id range [first, first + count) of `rank`; the first n_total % world ranks get one more."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, extra = divmod(int(n_total), int(world))
    count = base + (1 if rank < extra else 0)
    first = rank * base + min(rank, extra)
    return first, count


def allreduce_stats(stats_tensor, group=None):
    """In-place sum of the statistics block over all ranks; returns a dict (valid on every rank)."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(stats_tensor, op=dist.ReduceOp.SUM, group=group)
    vals = stats_tensor.detach().cpu().tolist()
    return dict(zip(abi.STATS_FIELDS, vals))


def summarize(stats):
    """Episode statistics in the form the reference's logging callback reports (scripts/train_flat_ppo.py:323-331)."""
    ep = max(stats["episodes"], 1.0)
    return {"episodes": stats["episodes"], "success_rate": stats["successes"] / ep, "mean_return": stats["return_sum"] / ep,
            "mean_length": stats["length_sum"] / ep, "mean_min_distance": stats["min_distance_sum"] / ep,
            "mean_final_distance": stats["final_distance_sum"] / ep, "env_steps": stats["env_steps"],
            "onboard_lock_fraction": stats["onboard_locks"] / max(stats["env_steps"], 1.0),
            "termination": {k: stats[k] for k in ("hit_target", "interceptor_crash", "fuel_out", "missile_ground",
                                                  "worsening", "timeouts")}}


class ShardedSim:
    """This rank's shard of an n_total-env simulation (device = LOCAL_RANK)."""

    def __init__(self, env_cfg, n_total, seed=1234, precision="fp32", rank=None, world=None, device=None, warn_dead=True):
        from .sim import HlynrSim

        r, w, lr = env_rank_world()
        self.rank = r if rank is None else rank
        self.world = w if world is None else world
        self.first, self.count = shard_range(n_total, self.rank, self.world)
        self.sim = HlynrSim(env_cfg, n_envs=self.count, device=lr if device is None else device, seed=seed,
                            env_id_offset=self.first, precision=precision, warn_dead=warn_dead)

    def rollout_stats(self, zero_after=True):
        """One all-reduce per rollout: global episode statistics."""
        t = self.sim.stats_tensor()
        out = allreduce_stats(t)
        if zero_after:
            self.sim.stats(zero_after=True)
        return out
