"""Spectral initialisation (reference stm.py:30-296).

CPU: the NumPy oracle (oracle/spectral_numpy.py) against tests/golden/spectral.npz — the LIVE
reference's gram / fastAnchor / spectral_init outputs (solve_qp shimmed by exact NNLS) on two synthetic
corpora and on the reference's shipped wiki corpus.
GPU (`-m gpu`): stm_spectral_gram / stm_spectral_finish through the C ABI against the same fixture and
against the oracle on seeded corpora; the reference's assertion behaviour; the STM front.

Tolerance: anchors EQUAL; beta <= 1e-8 relative to max(beta) — the per-word QP is solved by a different
(exact) active-set method than the reference's quadprog, so agreement is to solver rounding times the
conditioning of the anchor Gram matrix, not bit-exact (observed <= 1e-11).
"""
import numpy as np
import pytest

from conftest import load_golden, synthetic_corpus
from oracle import spectral_numpy as sn


def _case(g, tag):
    D, V, K, maxV = (int(x) for x in g[tag + "_cfg"][:4])
    V = g[tag + "_beta"].shape[1]
    return g[tag + "_doc_ptr"], g[tag + "_word_id"], g[tag + "_count"].astype(np.float64), K, V, maxV


@pytest.mark.parametrize("tag", ["t", "f"])
def test_oracle_stages_vs_live_reference(tag):
    g = load_golden("spectral.npz")
    ptr, ids, cnt, K, V, maxV = _case(g, tag)
    wprob = sn.word_prob(ptr, ids, cnt)
    keep = sn.keep_order(wprob, maxV)
    np.testing.assert_array_equal(keep, g[tag + "_keep"])
    Q = sn.gram(sn.dense_dtm(ptr, ids, cnt)[:, keep])
    np.testing.assert_allclose(Q, g[tag + "_Q"], rtol=1e-12, atol=1e-14)   # NOT row-normalised (stm.py:156 is a no-op)
    basis, _ = sn.fast_anchor(Q, K)
    np.testing.assert_array_equal(np.intp(basis), g[tag + "_anchor"])
    beta, anchors, _ = sn.spectral_init(ptr, ids, cnt, K, V, maxV)
    ref = g[tag + "_beta"]
    assert np.abs(beta - ref).max() <= 1e-9 * ref.max()
    np.testing.assert_allclose(beta.sum(axis=1), 1.0 / K, rtol=1e-9)   # total-sum normalisation, stm.py:82


@pytest.mark.parametrize("tag", ["t", "f"])
def test_oracle_fast_path_vs_live_reference(tag):
    """spectral_init_fast (sparse Gram + compiled NNLS: what bench.py's reference arm prepares the benchmark state
    with at D=100k, where the dense document-term matrix does not fit) against the live reference's fixture."""
    g = load_golden("spectral.npz")
    ptr, ids, cnt, K, V, maxV = _case(g, tag)
    keep = sn.keep_order(sn.word_prob(ptr, ids, cnt), maxV)
    np.testing.assert_allclose(sn.gram_sparse(ptr, ids, cnt, keep), g[tag + "_Q"], rtol=1e-11, atol=1e-14)
    beta, anchors, _ = sn.spectral_init_fast(ptr, ids, cnt, K, V, maxV)
    ref = g[tag + "_beta"]
    np.testing.assert_array_equal(anchors, keep[g[tag + "_anchor"]])
    assert np.abs(beta - ref).max() <= 1e-8 * ref.max()


def test_oracle_wiki_vs_live_reference():
    g, w = load_golden("spectral.npz"), load_golden("wiki_corpus.npz")
    K = int(g["w_cfg"][2])
    beta, anchors, keep = sn.spectral_init(w["doc_ptr"], w["word_id"], w["count"].astype(np.float64), K, int(w["V"]), 5000)
    np.testing.assert_array_equal(keep, g["w_keep"])
    np.testing.assert_array_equal(anchors, g["w_keep"][g["w_anchor"]])
    ref = g["w_beta_cols"]
    assert np.abs(beta[:, g["w_cols"]] - ref).max() <= 1e-9 * ref.max()
    np.testing.assert_allclose(beta.sum(axis=1), g["w_beta_rowsum"], rtol=1e-10)
    dropped = np.setdiff1d(np.arange(int(w["V"])), keep)[0]
    assert abs(beta[0, dropped] - float(g["w_beta_dropped"])) <= 1e-18


def test_nnls_is_the_qp_minimiser():
    """KKT conditions of min 1/2 w'Pw - q'w, w >= 0 on random strictly convex problems."""
    rng = np.random.default_rng(5)
    for K in (3, 10, 40):
        M = rng.random((K, 3 * K))
        P = M @ M.T
        for _ in range(5):
            q = M @ rng.normal(size=3 * K)
            w = sn.nnls_gram(P, q)
            g = q - P @ w
            assert np.all(w >= 0)
            assert np.all(g[w == 0] <= 1e-9 * np.abs(q).max())
            assert np.all(np.abs(g[w > 0]) <= 1e-9 * np.abs(q).max())


# ---- GPU -------------------------------------------------------------------------------------------

def _gpu_spectral(ptr, ids, cnt, K, V, maxV):
    from strutopy_b200.spectral import spectral_init
    return spectral_init((ptr, ids, cnt), K, V, maxV=maxV, verbose=False, return_anchors=True)


@pytest.mark.gpu
@pytest.mark.parametrize("tag", ["t", "f"])
def test_gpu_spectral_vs_live_reference_fixture(tag):
    g = load_golden("spectral.npz")
    ptr, ids, cnt, K, V, maxV = _case(g, tag)
    beta, anchors = _gpu_spectral(ptr, ids, cnt, K, V, maxV)
    np.testing.assert_array_equal(anchors, g[tag + "_keep"][g[tag + "_anchor"]])
    ref = g[tag + "_beta"]
    assert beta.shape == ref.shape
    assert np.abs(beta - ref).max() <= 1e-8 * ref.max()


@pytest.mark.gpu
def test_gpu_spectral_wiki_vs_live_reference():
    g, w = load_golden("spectral.npz"), load_golden("wiki_corpus.npz")
    K, V = int(g["w_cfg"][2]), int(w["V"])
    beta, anchors = _gpu_spectral(w["doc_ptr"], w["word_id"], w["count"].astype(np.float64), K, V, 5000)
    np.testing.assert_array_equal(anchors, g["w_keep"][g["w_anchor"]])
    ref = g["w_beta_cols"]
    assert np.abs(beta[:, g["w_cols"]] - ref).max() <= 1e-8 * ref.max()
    np.testing.assert_allclose(beta.sum(axis=1), g["w_beta_rowsum"], rtol=1e-9)


@pytest.mark.gpu
@pytest.mark.parametrize("D,V,K,maxV,nw", [(3000, 2500, 30, 1200, 100), (800, 4000, 60, 700, 150), (5000, 6000, 12, 5000, 120)])
def test_gpu_spectral_vs_oracle(D, V, K, maxV, nw):
    ptr, ids, cnt, _, _ = synthetic_corpus(D, V, K, n_words=nw, seed=D + K)
    if maxV >= V:   # keep everything: every word must occur (stm.py:152-154)
        seen = np.unique(ids)
        ids = np.searchsorted(seen, ids).astype(np.int32)
        V = len(seen)
    ref, ref_anchors, _ = sn.spectral_init(ptr, ids, cnt, K, V, maxV)
    beta, anchors = _gpu_spectral(ptr, ids, cnt, K, V, maxV)
    np.testing.assert_array_equal(anchors, ref_anchors)
    assert np.abs(beta - ref).max() <= 1e-8 * ref.max()
    assert np.all(beta > 0)


@pytest.mark.gpu
def test_gpu_spectral_reference_assertions():
    """stm.py:152-154: a never-occurring word among the kept ones, or a document with < 2 kept tokens,
    makes the reference's row-sum assertion fail."""
    ptr, ids, cnt, _, _ = synthetic_corpus(300, 400, 5, n_words=60, seed=7)   # some of the 400 words never occur
    with pytest.raises(AssertionError):
        _gpu_spectral(ptr, ids, cnt, 5, 400, 5000)
    ptr = np.array([0, 2, 3], np.int64)
    with pytest.raises(AssertionError):
        _gpu_spectral(ptr, np.array([0, 1, 1], np.int32), np.array([2.0, 1.0, 1.0]), 2, 2, 5000)


@pytest.mark.gpu
def test_front_spectral_init_and_fit():
    """STM(init_type='spectral') — the reference's default — initialises beta on the device and the fit runs."""
    from strutopy_b200 import STM
    g = load_golden("spectral.npz")
    ptr, ids, cnt, K, V, _ = _case(g, "f")
    D = len(ptr) - 1
    X = (np.arange(D) % 2).astype(np.float64)[:, None]
    m = STM((ptr, ids, cnt), range(V), False, K, X, False, 3, 0, 1e-5, init_type="spectral", model_type="STM")
    ref = g["f_beta"]
    assert np.abs(m.beta - ref).max() <= 2e-7 * ref.max()   # fp32 storage of beta on the device
    m.expectation_maximization(saving=False)
    assert len(m.last_bounds) == 3 and np.all(np.isfinite(m.last_bounds))
    assert m.last_bounds[-1] > m.last_bounds[0]


@pytest.mark.gpu
def test_wiki_corpus_from_mm_file_spectral_fit(tmp_path):
    """INTEGRATION.md's example: the reference's shipped corpus through an MmCorpus file, spectral init, a short fit."""
    from strutopy_b200 import STM
    from strutopy_b200.corpus import read_mm, write_mm
    g, w = load_golden("spectral.npz"), load_golden("wiki_corpus.npz")
    write_mm(tmp_path / "BoW_corpus.mm", w["doc_ptr"], w["word_id"], w["count"], int(w["V"]))
    ptr, ids, cnt, V = read_mm(tmp_path / "BoW_corpus.mm")
    K = int(g["w_cfg"][2])
    m = STM(documents=(ptr, ids, cnt), dictionary=range(V), X=w["X_50"], K=K, content=False, kappa_interactions=False,
            max_em_iter=3, sigma_prior=0, convergence_threshold=1e-5, lda_beta=True, init_type="spectral",
            model_type="STM")
    ref = g["w_beta_cols"]
    assert np.abs(m.beta[:, g["w_cols"]] - ref).max() <= 2e-7 * ref.max()      # fp32 storage of beta
    m.expectation_maximization(saving=False)
    assert len(m.last_bounds) == 3 and np.all(np.isfinite(m.last_bounds))
    prob, frex = m.label_topics(None, 5)
    assert len(prob) == K and all(0 <= i < V for i in prob[0])
