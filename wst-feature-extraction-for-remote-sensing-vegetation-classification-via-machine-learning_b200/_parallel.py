"""Multi-GPU sharding of the patch batch.

Every (patch, channel) signal is independent (the reference loops over them serially,
train_and_save_model.py:486,364), so the batch axis is split contiguously across ranks — one
process per GPU — with no data-path collective.  The only exchange is an optional all-gather of
the [B/G, F] feature shards over NCCL/NVLink so that every rank (or rank 0) holds the [B, F]
matrix in input order, which is what the Random-Forest trainer zips with its labels
(train_and_save_model.py:490,509-511).
"""
import torch
import torch.distributed as dist

__all__ = ["shard_range", "shard_sizes", "gather_features"]


def shard_range(B, rank, world_size):
    """Contiguous shard [lo, hi) of a batch of B patches for `rank` of `world_size`."""
    if world_size < 1 or not (0 <= rank < world_size):
        raise ValueError("invalid rank/world_size")
    return (rank * B) // world_size, ((rank + 1) * B) // world_size


def shard_sizes(B, world_size):
    return [shard_range(B, r, world_size)[1] - shard_range(B, r, world_size)[0] for r in range(world_size)]


def gather_features(local_feats, B, group=None, async_op=False, out=None):
    """All-gather row shards [b_r, F] into [B, F] in rank (= input) order.

    Works on whatever device `local_feats` lives on: NCCL for CUDA tensors, gloo for CPU tensors (the CPU form is what
    the world_size-2 tests exercise).  Equal shards (B divisible by the world size — the benchmark's case) are gathered
    straight into the output matrix by one all_gather_into_tensor: no padding copy, no concatenation.  Ragged shards
    are padded to the largest shard, gathered, and compacted.

    async_op=True (equal shards only) returns (out, work): the collective runs on the backend's own stream, so the
    caller can launch the next batch's kernels while the gather is in flight, and calls work.wait() before reading
    `out` (which may be passed in to reuse a buffer)."""
    if not dist.is_available() or not dist.is_initialized():
        if local_feats.shape[0] != B:
            raise ValueError("not distributed: local shard must be the whole batch")
        return (local_feats, None) if async_op else local_feats
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    sizes = shard_sizes(B, world)
    if local_feats.shape[0] != sizes[rank]:
        raise ValueError("rank %d holds %d rows, expected %d" % (rank, local_feats.shape[0], sizes[rank]))
    F = local_feats.shape[1]
    if min(sizes) == max(sizes):
        if out is None:
            out = local_feats.new_empty((B, F))
        elif tuple(out.shape) != (B, F) or not out.is_contiguous():
            raise ValueError("out must be a contiguous [B, F] tensor")
        work = dist.all_gather_into_tensor(out, local_feats.contiguous(), group=group, async_op=async_op)
        return (out, work) if async_op else out
    if async_op:
        raise ValueError("async_op needs equal shards (B divisible by the world size)")
    mx = max(sizes)
    padded = local_feats.new_zeros((mx, F))
    padded[: sizes[rank]] = local_feats
    buf = local_feats.new_empty((world * mx, F))
    dist.all_gather_into_tensor(buf, padded, group=group)
    buf = buf.view(world, mx, F)
    return torch.cat([buf[r, : sizes[r]] for r in range(world)], dim=0)
