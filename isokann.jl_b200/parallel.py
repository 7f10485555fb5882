"""Host-side sharding plan for one-process-per-GPU runs (nothing to mirror in
the reference, which is single-device).

Start points n (with their K Koopman samples) are split contiguously over ranks for the
featurize -> chi -> K-mean pass; every minibatch ``perm[i*B:(i+1)*B]`` is split contiguously
over ranks for the training step.  Both use the same rule as the library
(``split_range`` in csrc/api.cu); these helpers are what the host uses to slice ``ys`` and what
the gloo tests check.
"""
from __future__ import annotations

from typing import Optional, Tuple

import numpy as np


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """contiguous split of [0, n): returns (offset, length) of ``rank``; the first n % world
    ranks get one extra element."""
    base, rem = divmod(n, world)
    return rank * base + min(rank, rem), base + (1 if rank < rem else 0)


def batch_bounds(n: int, minibatch: int, partial: bool = False):
    """(start, length) of every minibatch of an epoch: batchsize rule of src/iso.jl:180 and the
    dropped tail of DataLoader(partial=false) (src/iso.jl:181)."""
    bs = n if (minibatch == 0 or n < minibatch) else minibatch
    nb = -(-n // bs) if partial else n // bs
    return [(i * bs, min(bs, n - i * bs)) for i in range(nb)]


def rank_batch_slice(perm1: np.ndarray, start: int, length: int, world: int, rank: int) -> np.ndarray:
    """the 1-based sample ids of minibatch [start, start+length) that ``rank`` processes"""
    off, n = shard_range(length, world, rank)
    return np.asarray(perm1[start + off:start + off + n])


def exchange_slices(lo: int, hi: int, world: int):
    """Who sums what in the gradient exchange over peer memory (csrc/p2p.cu): the flat range [lo, hi) is cut on
    4-element boundaries of the global index; rank r owns the r-th of `world` equal runs of float4 vectors, rank 0
    additionally the unaligned head and tail.  Returns per rank a list of (start, stop) element ranges."""
    lo4, hi4 = (lo + 3) & ~3, hi & ~3
    out = [[] for _ in range(world)]
    if hi4 > lo4:
        nvec = (hi4 - lo4) >> 2
        per = -(-nvec // world)
        for r in range(world):
            v0, v1 = min(nvec, per * r), min(nvec, per * r + per)
            if v1 > v0:
                out[r].append((lo4 + 4 * v0, lo4 + 4 * v1))
    head_end = min(lo4, hi)
    tail_start = max(hi4, head_end)
    if head_end > lo:
        out[0].append((lo, head_end))
    if hi > tail_start:
        out[0].append((tail_start, hi))
    return out


def reshard_plan(n_old: int, shift: int, n_new: int, world: int, rank: int):
    """Incremental data on several ranks (isokann_append_data: shift = 0, n_new = n_old + appended;
    isokann_keep_last: shift = dropped, n_new = kept): new start point i is old start point i + shift while that is
    < n_old, else row i + shift - n_old of the appended block.  Returns for ``rank``
    ((offset, length) of its new shard, (a, b) old global rows it keeps -> local rows [0, b - a),
    (a2, b2) "old-global" rows it takes from the appended block -> appended rows [a2 - n_old, b2 - n_old)),
    the rule of reshard_ys in csrc/api.cu."""
    o1, l1 = shard_range(n_new, world, rank)
    a, b = o1 + shift, min(o1 + l1 + shift, n_old)
    a2, b2 = max(o1 + shift, n_old), o1 + l1 + shift
    return (o1, l1), (a, max(a, b)), (a2, max(a2, b2))


def broadcast_unique_id(rank: int, src: int = 0) -> Optional[bytes]:
    """rank ``src`` creates the NCCL unique id, torch.distributed broadcasts it to everyone"""
    import torch.distributed as dist
    from .engine import Engine
    box = [Engine.unique_id() if rank == src else None]
    dist.broadcast_object_list(box, src=src)
    return box[0]
