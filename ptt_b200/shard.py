"""Batch-axis sharding of the hot path across the GPUs of one box.

Frames are independent in the forward pass, so rank r of W owns the contiguous slice
[r*B/W, (r+1)*B/W) of a global batch, weights are replicated, and the data path needs NO collective.
torch.distributed (NCCL on the GPU box, gloo in CPU tests) is used only for the barrier and the
max-over-ranks of the step time that bench.py reports.
"""
import torch
import torch.distributed as dist


def shard_range(total, rank, world):
    """Contiguous [lo, hi) of `total` frames owned by `rank`; the first total % world ranks get one extra frame."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank/world: %r/%r" % (rank, world))
    base, extra = divmod(int(total), world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def shard_batch(tensors, rank, world):
    """Slice every (B, ...) tensor of a dict (or a single tensor) to this rank's frames."""
    if isinstance(tensors, torch.Tensor):
        lo, hi = shard_range(tensors.shape[0], rank, world)
        return tensors[lo:hi]
    return {k: shard_batch(v, rank, world) for k, v in tensors.items()}


def max_over_ranks(values, device="cpu"):
    """Element-wise maximum of a list of floats over all ranks (identity without a process group)."""
    t = torch.tensor(list(values), dtype=torch.float64, device=device)
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return [float(x) for x in t.tolist()]


def whole_job_throughput(frames_per_rank, world, seconds_max):
    """Aggregate frames/s of a weak-scaling run: every rank processed `frames_per_rank` in at most `seconds_max`."""
    return frames_per_rank * world / seconds_max
