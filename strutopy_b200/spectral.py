"""Spectral initialisation of beta on the GPU — the reference's `spectral_init` (stm.py:30-84) with its
helpers `create_dtm` / `gram` / `fastAnchor` / `recover_l2` (stm.py:87-296).  SURVEY.md §8f-1.

Host side (this file, mirrors stm.py:51-59): word probabilities and the kept-word list
`np.argsort(-wprob)[:maxV]` — the same NumPy call as the reference, so ties break identically.
Device side (`stm_spectral_gram` / `stm_spectral_finish`, include/stm_b200.h): the Gram matrix
Q = Htilde'Htilde - Hhat (sparse: one warp per document adds its outer product into the packed triangle), the K anchor passes over Q, one
exact NNLS per word for recover_l2, and the K x V re-expansion.  With documents sharded over ranks the
local Gram statistics are all-reduced ONCE between the two calls; everything after is replicated.

The reference solves each word's QP with qpsolvers/quadprog; the QP is strictly convex, so the exact
active-set NNLS used here has the same (unique) minimiser — parity is to solver rounding, not bit-exact
(DESIGN.md §5).  There is no CPU fallback.
"""
import ctypes as C

import numpy as np

from . import _lib
from .corpus import pack_corpus
from .parallel import allreduce_stats


def keep_list(word_totals, maxV):
    """stm.py:53-59 -> (keep, wprob[keep]); `word_totals` are the column sums of the document-term matrix
    over ids 0..max id (create_dtm's width, stm.py:119)."""
    wprob = word_totals / np.sum(word_totals)
    keep = np.argsort(-1 * wprob)[:maxV]
    return keep, wprob[keep]


def spectral_on_context(ctx, torch, dev, word_totals, maxV=5000, dist=None, return_anchors=False):
    """spectral_init over the corpus resident on `ctx` (this rank's shard).  `word_totals`: GLOBAL column
    sums.  Returns beta (K x V host array) [, anchor word ids]."""
    L, h = _lib.load(), ctx.handle
    keep, wkeep = keep_list(np.asarray(word_totals, dtype=np.float64), maxV)
    n = int(keep.shape[0])
    keep32 = np.ascontiguousarray(keep, dtype=np.int32)
    wkeep = np.ascontiguousarray(wkeep, dtype=np.float64)
    st = torch.cuda.current_stream(dev).cuda_stream
    gram = torch.empty(n * n + n, dtype=torch.float64, device=dev)
    err = None
    try:
        _lib.check(h, L.stm_spectral_gram(h, n, _lib.hp(keep32), gram.data_ptr(), st))
    except _lib.StmError as e:   # raised after the collective below so that all ranks stay in step
        err = e
    if dist is not None:
        bad = torch.tensor([1 if err is not None else 0], dtype=torch.int32, device=dev)
        allreduce_stats(bad, dist)
        if int(bad.item()) and err is None:
            err = _lib.StmError(_lib.STM_ERR_INVALID, "Encountered zeroes in Q row sums, can not normalize. (another rank)")
        if err is None:
            # the packed upper triangle of Htilde'Htilde (n (n + 1) / 2 doubles: 100 MB at maxV = 5000) and diag(Hhat)
            allreduce_stats(gram[:n * (n + 1) // 2], dist)
            allreduce_stats(gram[n * n:], dist)
    if err is not None:
        _raise(err)
    beta = torch.empty((ctx.K, ctx.V), dtype=torch.float64, device=dev)
    anchors = np.zeros(ctx.K, dtype=np.int32)
    try:
        _lib.check(h, L.stm_spectral_finish(h, n, _lib.hp(keep32), _lib.hp(wkeep), gram.data_ptr(), beta.data_ptr(),
                                            _lib.hp(anchors), st))
    except _lib.StmError as e:
        _raise(e)
    b = beta.cpu().numpy()
    return (b, keep[anchors]) if return_anchors else b


def _raise(e):
    if e.code == _lib.STM_ERR_INVALID and "row sums" in str(e):
        raise AssertionError("Encountered zeroes in Q row sums, can not normalize.") from e   # stm.py:152-154
    raise e


def spectral_init(corpus, K, V, maxV=5000, verbose=True, print_anchor=False, device=0, return_anchors=False):
    """Same signature and result as the reference's spectral_init (stm.py:30-84): K x V beta whose rows sum
    to ~1/K (total-sum normalisation, stm.py:82).  `corpus`: list of [(word_id, count), ...] or a CSR triple."""
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("strutopy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    ptr, ids, cnt = pack_corpus(corpus)
    if ids.size == 0:
        raise ValueError("empty corpus")
    totals = np.bincount(ids, weights=cnt.astype(np.float64), minlength=int(ids.max()) + 1)
    dev = torch.device("cuda", int(device))
    ctx = _lib.Context(K, V, 1, int(device))
    try:
        ctx.set_corpus(ptr, ids, cnt)
        out = spectral_on_context(ctx, torch, dev, totals, maxV=maxV, return_anchors=True)
    finally:
        ctx.close()
    if print_anchor:
        for i, idx in enumerate(out[1]):
            print(f"{i}. anchor word: {int(idx)}")
    return out if return_anchors else out[0]
