"""Held-out likelihood by document completion — the reference's `modules/heldout.py` on the GPU.

Mirrors /root/reference/src/modules/heldout.py: `cut_in_half` (:70-85), `eval_heldout` (:88-97) and the
split used by `05_train.py:99-122`.  `eval_heldout` keeps the reference signature
(heldout, theta, beta) -> float; the arithmetic runs in `stm_heldout_host` (include/stm_b200.h): one warp per
document, fp64, no CPU fallback.  `STM.eval_heldout(heldout)` (strutopy_b200/stm.py) is the device-resident form:
it scores the fitted model's own theta / beta without copying them to the host."""
import numpy as np

from . import _lib
from .corpus import pack_corpus


def cut_in_half(doc_set):
    """Every other (word, count) pair of each document: (docs[0::2], docs[1::2])   heldout.py:70-85"""
    first_half = np.zeros(len(doc_set), dtype=np.ndarray)
    second_half = np.zeros(len(doc_set), dtype=np.ndarray)
    for i in range(len(doc_set)):
        first_half[i] = doc_set[i][0::2]
        second_half[i] = doc_set[i][1::2]
    return first_half, second_half


def split_corpus(corpus, validation_set=False, document_completion=True, proportion=0.8):
    """Train / test split by position (heldout.py:40-67).  The reference leaves `validate_docs` unbound when
    `validation_set` is False and `test_1_docs` / `test_2_docs` unbound without document completion (NameError);
    here those come back as None."""
    corpus = [doc for doc in corpus]
    test_split_idx = int(proportion * len(corpus))
    train_docs = corpus[:test_split_idx]
    validate_docs = None
    if validation_set:
        validate_split_idx = int((proportion + (1 - proportion) / 2) * len(corpus))
        test_docs = corpus[test_split_idx:validate_split_idx]
        validate_docs = corpus[validate_split_idx:]
    else:
        test_docs = corpus[test_split_idx:]
    test_1_docs = test_2_docs = None
    if document_completion:
        test_1_docs, test_2_docs = cut_in_half(test_docs)
    return train_docs, test_1_docs, test_2_docs, validate_docs


def eval_heldout(heldout, theta, beta, device=0, return_doc_ll=False):
    """mean over documents of  sum_w c_w log(theta_i . beta[:, w]) / sum_w c_w      heldout.py:88-97

    heldout: list of documents [(word_id, count), ...] (document i is scored with theta[i]) or a CSR triple
    (doc_ptr, word_id, count); theta: (D, K); beta: (K, V)."""
    theta = np.asarray(theta, dtype=np.float64)
    beta = np.asarray(beta, dtype=np.float64)
    if beta.ndim != 2 or theta.ndim != 2 or theta.shape[1] != beta.shape[0]:
        raise ValueError("theta must be (D, K) and beta (K, V)")
    ptr, ids, cnt = pack_corpus(list(heldout) if not isinstance(heldout, tuple) else heldout)
    D = ptr.shape[0] - 1
    if D < 1:
        raise ValueError("no held-out documents")
    if theta.shape[0] < D:
        raise IndexError("fewer rows in theta than held-out documents")   # the reference's theta[i] would raise
    if ids.size and (ids.min() < -beta.shape[1] or ids.max() >= beta.shape[1]):
        raise IndexError("word id out of range for beta")                 # the reference's beta[:, w] would raise
    if ids.size and ids.min() < 0:
        ids = np.where(ids < 0, ids + beta.shape[1], ids).astype(np.int32)   # NumPy's negative indexing
    ctx = _lib.Context(beta.shape[0], beta.shape[1], 1, device)
    try:
        mean, doc_ll = ctx.heldout_host(ptr, ids, cnt, theta[:D], beta)
    finally:
        ctx.close()
    return (mean, doc_ll) if return_doc_ll else mean
