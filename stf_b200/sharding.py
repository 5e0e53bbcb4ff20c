"""Batch sharding across the GPUs of one box: images are independent, so rank r of W
codes the contiguous block of images [r*k, (r+1)*k) with no data-path collective; only byte strings / timings
are gathered for reporting.  Works with any torch.distributed backend (nccl on the GPU box, gloo in the tests)."""
import torch
import torch.distributed as dist


def shard_range(n_items, rank, world):
    """Contiguous block of `n_items` owned by `rank` (blocks differ by at most one item)."""
    if not 0 <= rank < world:
        raise ValueError(f"rank {rank} outside world of {world}")
    base, extra = divmod(n_items, world)
    start = rank * base + min(rank, extra)
    return start, start + base + (1 if rank < extra else 0)


def gather_strings(local_strings, group=None):
    """All ranks' per-image byte strings in global image order (every rank gets the full list)."""
    if not dist.is_available() or not dist.is_initialized():
        return list(local_strings)
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, list(local_strings), group=group)
    return [s for part in out for s in part]


def max_over_ranks(value, device=None, group=None):
    """max of a host scalar over all ranks (multi-GPU timings are the slowest rank's)."""
    if not dist.is_available() or not dist.is_initialized():
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
