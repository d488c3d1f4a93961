"""Host-side sharding of an environment population across ranks (one process per GPU).

Environments are independent, so the per-step path has no collective: rank r of
G owns the contiguous global range [offset, offset + count), passes `env_offset=offset` to its
BatchedCookingEnv (so layout draws are keyed by the *global* environment index and results do
not depend on G), and episode statistics are summed with one all_reduce when somebody asks.
"""
import numpy as np
import torch
import torch.distributed as dist


def shard_range(total_envs, world_size, rank):
    """Contiguous split; the first `total % world` ranks get one extra environment."""
    if not 0 <= rank < world_size:
        raise ValueError("rank out of range")
    base, extra = divmod(int(total_envs), int(world_size))
    count = base + (1 if rank < extra else 0)
    offset = rank * base + min(rank, extra)
    return offset, count


def global_layout_ids(draw, seed, offset, count, num_layouts, episode=0):
    """Initial layout id of every environment of a shard; `draw` is lib.cz_layout_draw."""
    return np.array([draw(seed, offset + e, episode) % num_layouts for e in range(count)], np.int32)


def shard_rows(array, world_size, rank):
    """Rows of a global per-environment array (e.g. recipe ids) that belong to `rank`."""
    offset, count = shard_range(len(array), world_size, rank)
    return array[offset:offset + count]


def reduce_stats(stats, group=None):
    """Sum a small vector of episode statistics over all ranks (NCCL on GPUs, gloo on CPU)."""
    t = stats if isinstance(stats, torch.Tensor) else torch.as_tensor(stats, dtype=torch.float64)
    t = t.clone()
    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.SUM, group=group)
    return t
