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


def allreduce_stats(stats, dist=None):
    """THE collective of the data-parallel path: a sum-all-reduce, in place, of a torch tensor (CUDA with NCCL on the
    GPU box, CPU with gloo in the tests).  Every exchange of the product goes through here: the packed statistics
    buffer once per EM iteration (STM._reduce_and_bound), the moments in M_step(), the Gram statistics of the
    spectral initialisation, document / word totals in the constructor.  `dist` None (single process): no-op."""
    if dist is not None and dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(stats, op=dist.ReduceOp.SUM)
    return stats
