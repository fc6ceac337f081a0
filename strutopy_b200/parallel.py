"""Data-parallel plumbing: documents are sharded across ranks (one process per GPU); the only
exchange per EM iteration is ONE sum-all-reduce of the packed sufficient-statistics buffer
(include/stm_b200.h "packed sufficient-statistics buffer").  The reference has no
distributed path (its documents loop is serial, stm.py:519) — this module is new.

Host-only logic (NumPy + torch.distributed); runs on CPU with the gloo backend in the tests.
"""
import numpy as np


def shard_bounds(doc_ptr, world):
    """Contiguous document ranges [(lo, hi)] * world, balanced by nnz (per-document cost is
    proportional to the number of distinct words), never splitting a document."""
    doc_ptr = np.asarray(doc_ptr, dtype=np.int64)
    D = doc_ptr.shape[0] - 1
    if world <= 1:
        return [(0, D)]
    nnz = int(doc_ptr[-1])
    # weight = nnz + a per-document constant so that empty / tiny documents still spread out
    w = (doc_ptr[1:] - doc_ptr[:-1]).astype(np.float64) + (nnz / max(D, 1)) * 0.25 + 1e-9
    cum = np.concatenate([[0.0], np.cumsum(w)])
    cuts = [0]
    for r in range(1, world):
        target = cum[-1] * r / world
        c = int(np.searchsorted(cum, target, side="left"))
        c = min(max(c, cuts[-1]), D)
        cuts.append(c)
    cuts.append(D)
    return [(cuts[r], cuts[r + 1]) for r in range(world)]


def stats_layout(A, V, TS, K, p):
    """Python mirror of stm_stats_layout (include/stm_b200.h): segment offsets + total length."""
    K1 = K - 1
    sizes = [A * V * TS, K1 * K1, 1, 1, K1, p, p * p, p * K1, K1 * K1]
    off = [0]
    for s in sizes:
        off.append(off[-1] + s)
    return off


def allreduce_stats(stats, dist=None):
    """The one collective of an EM iteration.  `stats` is a torch tensor (CUDA with NCCL, CPU with
    gloo); reduced in place and returned."""
    if dist is None:
        import torch.distributed as dist  # noqa: PLW0642
    if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
