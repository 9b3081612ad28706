"""
Multi-GPU plumbing: one process per GPU (torchrun), reads sharded embarrassingly, result records gathered to the
writer rank.  The alignments themselves never communicate; the only exchange step is the gather
of fixed-size 9 x int32 records (36 B per alignment) -- over NCCL/NVLink on the GPU box, over gloo in the CPU tests.
"""
import numpy as np


def shard_bounds(n_items, world_size):
    """Contiguous, balanced-by-count shard boundaries: int64[world_size + 1]."""
    base, rem = divmod(int(n_items), int(world_size))
    sizes = [base + (1 if r < rem else 0) for r in range(world_size)]
    return np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)


def shard_bounds_by_bases(seq_off, world_size):
    """Contiguous shard boundaries balanced on the sum of sequence lengths (full-read scans)."""
    seq_off = np.asarray(seq_off, dtype=np.int64)
    n = len(seq_off) - 1
    total = int(seq_off[-1] - seq_off[0])
    bounds = [0]
    for r in range(1, world_size):
        target = seq_off[0] + (total * r) // world_size
        k = int(np.searchsorted(seq_off, target, side='left'))
        bounds.append(min(max(k, bounds[-1]), n))
    bounds.append(n)
    return np.asarray(bounds, dtype=np.int64)


def gather_records(local, counts, dst=0, group=None):
    """
    Gather per-rank record tensors (int32[count_r, 9], on the device for NCCL or on the CPU for gloo) to `dst`.
    `counts` are the per-rank record counts (known to every rank from the shard bounds).  Returns the
    concatenated int32[sum(counts), 9] tensor on dst, None elsewhere.  Ranks are padded to the largest shard so
    a single equal-size all_gather suffices (the volume is tiny next to the DP work).
    """
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    rank = dist.get_rank(group)
    if world == 1:
        return local
    mx = int(max(counts))
    pad = torch.zeros((mx, local.shape[1]), dtype=local.dtype, device=local.device)
    pad[:local.shape[0]] = local
    bucket = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bucket, pad, group=group)
    if rank != dst:
        return None
    return torch.cat([bucket[r][:int(counts[r])] for r in range(world)], dim=0)
