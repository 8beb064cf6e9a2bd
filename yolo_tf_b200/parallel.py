"""Data-parallel plumbing (one process per GPU, torch.distributed; NCCL on B200, gloo in CPU tests).

The path shards by image: inference + decode + NMS need NO exchange step; the
training step needs exactly one all-reduce(sum) on a flat float32 gradient bucket followed by 1/G
(the loss's cnt = B*cells*A uses the local batch, model/yolo2/__init__.py:89, so averaging replicas
reproduces the global-batch mean).  The reference has no multi-GPU support at all (README.md:99).
"""
import os


def world():
    """(rank, world_size, local_rank) from the torchrun environment (1-process defaults)."""
    return int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1")), int(os.environ.get("LOCAL_RANK", "0"))


def shard_range(total, rank, world_size):
    """Contiguous [begin, end) image range of `rank`; remainders go to the lowest ranks."""
    if not (0 <= rank < world_size):
        raise ValueError("rank %d out of range for world size %d" % (rank, world_size))
    base, rem = divmod(total, world_size)
    begin = rank * base + min(rank, rem)
    return begin, begin + base + (1 if rank < rem else 0)


def shard_batch(batch, rank, world_size):
    """Slice of a [B, ...] array/tensor owned by `rank` (images are independent units)."""
    b, e = shard_range(len(batch), rank, world_size)
    return batch[b:e]


def allreduce_mean_(flat_bucket, group=None):
    """In-place mean over replicas of one flat gradient bucket: ONE collective per training step."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return flat_bucket
    dist.all_reduce(flat_bucket, op=dist.ReduceOp.SUM, group=group)
    flat_bucket.div_(dist.get_world_size(group))
    return flat_bucket


def gather_detections(local, group=None):
    """Optional: collect each rank's (small) detection list on every rank, in rank order."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return [local]
    out = [None] * dist.get_world_size(group)
    dist.all_gather_object(out, local, group=group)
    return out


def max_over_ranks(value, device=None, group=None):
    """Timing rule: a multi-GPU number is the max over ranks, measured on the device."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX, group=group)
    return float(t.item())
