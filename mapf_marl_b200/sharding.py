"""Environments sharded by index over the GPUs of one box.

The path has no data-path exchange: environments never interact (each reference env instance is
self-contained, and the reference's own parallelism is one env per worker process,
MARL-curve-main/src/runners/parallel_runner.py:23-31).  Rank r owns the contiguous global range
[lo, hi); inputs are generated from the GLOBAL environment index, so results do not depend on the number
of GPUs.  The only collective is one all-reduce(SUM) of the 8-entry statistics vector at logging cadence
(the reference's equivalent is the runner's `_log` averaging, episode_runner.py:139-147): NCCL over
NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist

from ._lib import STAT_NAMES


def shard_range(n_envs_total, rank, world_size):
    """Contiguous, balanced split: the first (n % world) ranks own one extra environment."""
    base, extra = divmod(int(n_envs_total), int(world_size))
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def stats_to_tensor(stats, device):
    return torch.tensor([int(stats[k]) for k in STAT_NAMES], dtype=torch.int64, device=device)


def reduce_stats(stats, device=None, group=None):
    """Sum a rank's statistics dict over all ranks (no-op without an initialised process group)."""
    if device is None:
        device = torch.device("cuda", torch.cuda.current_device()) if torch.cuda.is_available() else torch.device("cpu")
    vec = stats_to_tensor(stats, device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(vec, op=dist.ReduceOp.SUM, group=group)
    return dict(zip(STAT_NAMES, [int(v) for v in vec.tolist()]))


class ShardedEngine:
    """One MapfEngine per rank over its slice of a global batch of environments."""

    def __init__(self, n_envs_total, rank=None, world_size=None, device=None, **engine_kwargs):
        from .engine import MapfEngine
        if rank is None:
            rank = dist.get_rank() if dist.is_initialized() else 0
        if world_size is None:
            world_size = dist.get_world_size() if dist.is_initialized() else 1
        self.rank, self.world_size = rank, world_size
        self.lo, self.hi = shard_range(n_envs_total, rank, world_size)
        self.engine = MapfEngine(self.hi - self.lo, device=device, **engine_kwargs)

    def global_stats(self):
        return reduce_stats(self.engine.stats(), self.engine.device)
