"""Batch sharding across the GPUs of one box.

Every polynomial / mat-vec instance / sampler stream is independent, so the batch index range [0, B) is
split into `world` contiguous slabs, rank g takes rows [lo, hi) of every operand tensor, and no data-path
collective exists.  torch.distributed is used only to agree on the timing (barrier + max over ranks) and,
for host-side consumers, to gather per-rank digests or results.
"""
import hashlib

import numpy as np


def shard_range(total, world, rank):
    """Contiguous, balanced split: the first `total % world` ranks take one extra row."""
    if world < 1 or not (0 <= rank < world):
        raise ValueError("bad rank %d of %d" % (rank, world))
    base, extra = divmod(total, world)
    lo = rank * base + min(rank, extra)
    hi = lo + base + (1 if rank < extra else 0)
    return lo, hi


def max_over_ranks(value, device=None):
    """Max of a scalar over all ranks (device time of the slowest rank = job time)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def sum_over_ranks(value, device=None):
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return float(value)
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return float(t.item())


def digest_rows(array):
    """Order-sensitive digest of a [rows, n] int32 slab (host side)."""
    return hashlib.sha256(np.ascontiguousarray(array).tobytes()).digest()


def gather_digests(local_digest):
    """Host-side gather of per-rank digests; the job digest is the digest of their concatenation in rank
    order, which equals the single-process 'checksum of checksums' over the same slabs."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size() == 1:
        return hashlib.sha256(local_digest).digest()
    out = [None] * dist.get_world_size()
    dist.all_gather_object(out, local_digest)
    return hashlib.sha256(b"".join(out)).digest()
