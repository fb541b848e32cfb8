"""Multi-GPU plumbing: environments are independent, so they shard across ranks with NO collective on the step
path; the only exchange is the reduction of the finished-episode statistics."""
import torch
import torch.distributed as dist


def shard_range(total_envs, rank, world_size):
    """Contiguous shard [lo, hi) of `total_envs` global env ids owned by `rank` (ids key the Philox streams, so
    results do not depend on the sharding)."""
    base, rem = divmod(int(total_envs), int(world_size))
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def reduce_episode_stats(stats, group=None):
    """Sum the (10,) int64 episode-stat vector over all ranks (NCCL on GPUs, gloo on CPU tensors)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return stats
    out = stats.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def reduce_return_sums(sums, group=None):
    """Sum the (2,) float64 [imposter, crew] return sums over all ranks (entries 10 and 11 of the episode-stat vector)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return sums
    out = sums.clone()
    dist.all_reduce(out, op=dist.ReduceOp.SUM, group=group)
    return out


def max_over_ranks(value, device=None, group=None):
    """Max of a python float over ranks (bench timing contract)."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
