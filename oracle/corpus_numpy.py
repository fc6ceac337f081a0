"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for the device corpus sampler
stm_sample_corpus (strutopy_b200/csrc/corpus_gen.cuh) — SURVEY.md §8f-3.

The reference samples documents with NumPy's PCG64 multinomial over the dense theta @ beta
(/root/reference/src/modules/generate_docs.py:293-316); a parallel sampler cannot reproduce that stream.
What is pinned here instead:
  * `sample_corpus`: a bit-exact NumPy restatement of the device sampler (Philox4x32-10 of Salmon et al.,
    SC'11, counter = (token, doc_lo, doc_hi, 0), key = seed; topic by inverse CDF of theta_d, word by
    inverse CDF of beta_z; per-document sort + run-length encoding) — the GPU corpus must EQUAL it;
  * `expected_word_mass`: sum_d theta_d beta, the mean of the reference's multinomial, for the
    distributional tests (tests/test_corpus_gen.py), which the reference's own sampler is held to as well.
"""
import numpy as np

M0, M1 = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57)
W0, W1 = 0x9E3779B9, 0xBB67AE85
MASK = np.uint64(0xFFFFFFFF)


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """vectorised over the counter words (uint64 arrays holding 32-bit values)"""
    c0, c1, c2, c3 = (np.asarray(x, dtype=np.uint64) for x in (c0, c1, c2, c3))
    k0, k1 = int(k0), int(k1)
    for _ in range(10):
        p0, p1 = M0 * c0, M1 * c2
        hi0, lo0, hi1, lo1 = p0 >> np.uint64(32), p0 & MASK, p1 >> np.uint64(32), p1 & MASK
        c0, c1, c2, c3 = hi1 ^ c1 ^ np.uint64(k0), lo1, hi0 ^ c3 ^ np.uint64(k1), lo0
        k0, k1 = (k0 + W0) & 0xFFFFFFFF, (k1 + W1) & 0xFFFFFFFF
    return c0, c1, c2, c3


def _u53(hi, lo):
    return (((hi << np.uint64(32)) | lo) >> np.uint64(11)).astype(np.float64) * 1.1102230246251565e-16


def sample_corpus(theta, beta, n_words, seed):
    """-> (doc_ptr int64, word_id int32, count float32), ids ascending within a document"""
    theta, beta = np.asarray(theta, np.float64), np.asarray(beta, np.float64)
    D, K = theta.shape
    V = beta.shape[1]
    cth = np.cumsum(theta, axis=1)
    cb = np.cumsum(beta, axis=1)
    ptr, ids, cnt = [0], [], []
    t = np.arange(n_words, dtype=np.uint64)
    for d in range(D):
        r0, r1, r2, r3 = philox4x32_10(t, np.full(n_words, d & 0xFFFFFFFF, np.uint64),
                                       np.full(n_words, d >> 32, np.uint64), np.zeros(n_words, np.uint64),
                                       seed & 0xFFFFFFFF, seed >> 32)
        z = np.minimum(np.searchsorted(cth[d], _u53(r0, r1) * cth[d, K - 1], side="right"), K - 1)
        w = np.empty(n_words, np.int64)
        u2 = _u53(r2, r3)
        for k in np.unique(z):
            sel = z == k
            w[sel] = np.minimum(np.searchsorted(cb[k], u2[sel] * cb[k, V - 1], side="right"), V - 1)
        u, c = np.unique(w, return_counts=True)
        ids.append(u.astype(np.int32))
        cnt.append(c.astype(np.float32))
        ptr.append(ptr[-1] + len(u))
    return np.array(ptr, np.int64), np.concatenate(ids), np.concatenate(cnt)


def expected_word_mass(theta, beta, n_words):
    """E[column sums of the document-term matrix] = n_words * sum_d theta_d beta (generate_docs.py:297-302)"""
    return n_words * (np.asarray(theta).sum(axis=0) @ np.asarray(beta))
