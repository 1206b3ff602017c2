"""Pair-wise sharding of a batch of alignments over the GPUs of one box, and the single collective
of the path: a gather of the fixed 256-byte result records.

Every (reference, current, guess) alignment is independent -- PwnCloser::processPartition loops
serially over independent candidates (pwn_tracker2/pwn_closer.cpp:92-105) -- so pairs are split
into contiguous blocks, one per rank, with no data-path collective; the only exchange is the
final gather of transforms and inlier statistics (NCCL over NVLink on GPUs, gloo in CPU tests).
"""
import numpy as np

RECORD_BYTES = 256


def partition(n_pairs, world_size, rank):
    """Contiguous block [lo, hi) of rank: GPU g gets pairs [g*n/G, (g+1)*n/G)."""
    lo = (n_pairs * rank) // world_size
    hi = (n_pairs * (rank + 1)) // world_size
    return lo, hi


def order_pairs_by_current(pairs):
    """Sort (reference_id, current_id) pairs so that pairs sharing a current cloud are contiguous
    (its z-buffer / index image is then computed once per chunk).  Returns the permutation."""
    pairs = np.asarray(pairs)
    return np.lexsort((pairs[:, 0], pairs[:, 1]))


def gather_records(local_records, n_total, device=None, group=None):
    """All-gather the per-rank record arrays (numpy structured arrays of 256-byte records, possibly of
    different lengths) into one array of n_total records ordered by rank.  Uses torch.distributed
    (NCCL for CUDA tensors, gloo for CPU tensors); with no initialised process group it is the
    identity."""
    import torch
    import torch.distributed as dist
    local_records = np.ascontiguousarray(local_records)
    assert local_records.dtype.itemsize == RECORD_BYTES
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        assert len(local_records) == n_total
        return local_records
    world = dist.get_world_size(group)
    counts = [partition(n_total, world, r) for r in range(world)]
    max_n = max(hi - lo for lo, hi in counts)
    buf = np.zeros((max_n, RECORD_BYTES), np.uint8)
    buf[:len(local_records)] = local_records.view(np.uint8).reshape(-1, RECORD_BYTES)
    t = torch.from_numpy(buf)
    if device is not None:
        t = t.to(device, non_blocking=True)
    out = torch.empty((world * max_n, RECORD_BYTES), dtype=torch.uint8, device=t.device)
    dist.all_gather_into_tensor(out, t, group=group)
    out = out.cpu().numpy().reshape(world, max_n, RECORD_BYTES)
    parts = [out[r, :hi - lo] for r, (lo, hi) in enumerate(counts)]
    return np.concatenate(parts, axis=0).reshape(-1).view(local_records.dtype)
