"""Synthetic corpus sampler (reference generate_docs.py:99-417).

CPU: the oracle's Philox4x32-10 against the Random123 known-answer vectors; the oracle sampler's
invariants and its word totals against n_words * sum_d theta_d beta (the mean of the reference's
multinomial, generate_docs.py:297-302).
GPU (`-m gpu`): stm_sample_corpus EQUALS the oracle (bit-exact: ids, counts, doc_ptr); edge sizes; the
`CorpusCreation` mirror passes the reference's own generate_docs tests (tests/test_generate_docs.py there).
"""
import numpy as np
import pytest

from oracle import corpus_numpy as cn


def _params(D, K, V, seed):
    rng = np.random.default_rng(seed)
    return rng.dirichlet(np.ones(K), D), rng.dirichlet(np.full(V, 0.05), K)


def test_philox_known_answers():
    """Random123 kat_vectors, philox4x32 with 10 rounds"""
    kat = [((0, 0, 0, 0), (0, 0), (0x6627e8d5, 0xe169c58d, 0xbc57ac4c, 0x9b00dbd8)),
           ((0xffffffff,) * 4, (0xffffffff,) * 2, (0x408f276d, 0x41c83b0e, 0xa20bc7c6, 0x6d5451fd)),
           ((0x243f6a88, 0x85a308d3, 0x13198a2e, 0x03707344), (0xa4093822, 0x299f31d0),
            (0xd16cfe09, 0x94fdcceb, 0x5001e420, 0x24126ea1))]
    for ctr, key, want in kat:
        got = cn.philox4x32_10(*[[c] for c in ctr], *key)
        assert tuple(int(g[0]) for g in got) == want


def _check_invariants(ptr, ids, cnt, D, V, n_words):
    assert ptr[0] == 0 and len(ptr) == D + 1 and ptr[-1] == len(ids) == len(cnt)
    assert ids.min() >= 0 and ids.max() < V
    for d in range(D):
        w = ids[ptr[d]:ptr[d + 1]]
        assert np.all(np.diff(w) > 0)                     # ascending, unique
    np.testing.assert_array_equal(np.add.reduceat(cnt, ptr[:-1]), n_words)
    assert np.all(cnt >= 1)


def test_oracle_sampler_distribution():
    D, K, V, n_words = 1500, 6, 200, 80
    theta, beta = _params(D, K, V, 3)
    ptr, ids, cnt = cn.sample_corpus(theta, beta, n_words, seed=99)
    _check_invariants(ptr, ids, cnt, D, V, n_words)
    tot = np.bincount(ids, weights=cnt, minlength=V)
    exp = cn.expected_word_mass(theta, beta, n_words)
    z = (tot - exp) / np.sqrt(np.maximum(exp, 1.0))
    assert np.abs(z[exp > 5]).max() < 5.0                 # Poisson-like fluctuations only
    assert abs(z[exp > 5].std() - 1.0) < 0.2
    # a different seed gives a different corpus, the same seed the same one
    p2, i2, c2 = cn.sample_corpus(theta[:50], beta, n_words, seed=99)
    np.testing.assert_array_equal(i2, ids[:ptr[50]])
    p3, i3, c3 = cn.sample_corpus(theta[:50], beta, n_words, seed=100)
    assert len(i3) != len(i2) or np.any(i3 != i2)


# ---- GPU -------------------------------------------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("D,K,V,n_words,seed", [(300, 7, 500, 150, 12345), (64, 50, 3000, 1, 7), (5, 3, 40, 4096, 2 ** 40 + 3),
                                                 (40, 128, 900, 333, 0)])
def test_gpu_sampler_equals_oracle(D, K, V, n_words, seed):
    from strutopy_b200.generate_docs import sample_corpus
    theta, beta = _params(D, K, V, D + K)
    ptr, ids, cnt = sample_corpus(theta, beta, n_words, seed=seed)
    rp, ri, rc = cn.sample_corpus(theta, beta, n_words, seed=seed)
    np.testing.assert_array_equal(ptr, rp)
    np.testing.assert_array_equal(ids, ri)
    np.testing.assert_array_equal(cnt, rc)
    _check_invariants(ptr, ids, cnt, D, V, n_words)


@pytest.mark.gpu
def test_gpu_sampler_distribution_config_shape():
    """20k documents of BASELINE config 3's shape: word totals follow n_words * sum_d theta_d beta."""
    from strutopy_b200.generate_docs import sample_corpus
    D, K, V, n_words = 20000, 50, 10000, 150
    theta, beta = _params(D, K, V, 11)
    ptr, ids, cnt = sample_corpus(theta, beta, n_words, seed=5)
    np.testing.assert_array_equal(np.add.reduceat(cnt, ptr[:-1]), n_words)
    tot = np.bincount(ids, weights=cnt, minlength=V)
    exp = cn.expected_word_mass(theta, beta, n_words)
    z = (tot - exp)[exp > 5] / np.sqrt(exp[exp > 5])
    assert np.abs(z).max() < 6.0 and abs(z.std() - 1.0) < 0.1


@pytest.mark.gpu
def test_corpus_creation_mirrors_reference_tests():
    """the reference's tests/test_generate_docs.py + conftest toy corpus, on the mirror class"""
    from strutopy_b200.generate_docs import CorpusCreation
    np.random.seed(42)
    K, N, n_words, V, level = 3, 50, 50, 200, 1
    gamma = np.random.multivariate_normal(np.random.standard_normal(level), np.diag(np.full(level, 0.001)), K - 1)
    corpus = CorpusCreation(n_topics=K, n_docs=N, n_words=n_words, V=V, level=level, gamma=gamma, dgp="STM")
    corpus.generate_documents(remove_terms=True)
    corpus.split_corpus(proportion=0.8)
    assert len(corpus.documents) == N
    assert corpus.theta.shape == (N, K) and corpus.beta.shape == (K, 200)
    np.testing.assert_allclose(corpus.theta.sum(axis=1), 1.0, atol=1e-10)
    np.testing.assert_allclose(corpus.beta.sum(axis=1), 1.0, atol=1e-10)
    for doc in corpus.documents:
        assert isinstance(doc, list) and sum(c for _, c in doc) == n_words
        for w, c in doc:
            assert isinstance(w, int) and c > 0 and 0 <= w < corpus.V
    assert corpus.V == len(corpus.dictionary) <= 200
    assert len(corpus.train_docs) == 40 and len(corpus.test_docs) == 10
    assert len(corpus.test_1_docs) == len(corpus.test_2_docs) == 10
    # first-appearance numbering (generate_docs.py:299-315) before the compaction: ids 0..n-1 all used
    c2 = CorpusCreation(n_topics=K, n_docs=N, n_words=n_words, V=V, level=level, gamma=gamma, dgp="LDA")
    c2.generate_documents(remove_terms=False)
    ids = c2.csr[1]
    first = [ids[i] for i in sorted(np.unique(ids, return_index=True)[1])]
    assert first == list(range(len(first)))
    # and the corpus feeds the STM front
    from strutopy_b200 import STM
    m = STM(corpus.csr, corpus.dictionary, False, K, corpus.metadata, False, 2, 0, 1e-5, init_type="random", model_type="STM")
    m.expectation_maximization(saving=False)
    assert np.all(np.isfinite(m.last_bounds))


def test_philox_uniformity():
    """53-bit uniforms from the oracle's Philox stream: mean / variance / serial correlation of 200k draws."""
    n = 200000
    t = np.arange(n, dtype=np.uint64)
    r0, r1, r2, r3 = cn.philox4x32_10(t & np.uint64(0xFFFFFFFF), np.full(n, 7, np.uint64), np.zeros(n, np.uint64),
                                      np.zeros(n, np.uint64), 12345, 0)
    for u in (cn._u53(r0, r1), cn._u53(r2, r3)):
        assert 0.0 <= u.min() and u.max() < 1.0
        assert abs(u.mean() - 0.5) < 4 * np.sqrt(1 / 12 / n)
        assert abs(u.var() - 1 / 12) < 1e-3
        assert abs(np.corrcoef(u[:-1], u[1:])[0, 1]) < 0.01
        hist = np.bincount((u * 64).astype(int), minlength=64)
        chi2 = ((hist - n / 64) ** 2 / (n / 64)).sum()
        assert chi2 < 120          # 63 degrees of freedom: P(chi2 > 120) ~ 1e-5


def test_corpus_creation_parameter_draws_equal_the_live_reference():
    """The constructor of the CorpusCreation mirror runs on the host: with np.random seeded as the reference's scripts
    seed it (04_create_synthetic_corpora.py:45-47) it reproduces the reference's alpha, beta, gamma, metadata, eta and
    theta bit for bit (fixture from the live reference, generate_docs.py:99-271)."""
    from conftest import load_golden
    from strutopy_b200.generate_docs import CorpusCreation
    g = load_golden("corpus_params.npz")
    cfgs = dict(stm=dict(dgp="STM", level=2), lda=dict(dgp="LDA", level=1),
                ldat=dict(dgp="LDA", level=1, treatment=True, alpha_treatment="auto-linear", alpha="asymmetric"))
    for tag, kw in cfgs.items():
        np.random.seed(12345)
        c = CorpusCreation(n_topics=5, n_docs=40, n_words=30, V=60, **kw)
        for k in ("alpha", "beta", "gamma", "metadata", "eta"):
            np.testing.assert_array_equal(np.asarray(getattr(c, k)), g[f"{tag}_{k}"], err_msg=f"{tag} {k}")
        np.testing.assert_allclose(c.theta, g[f"{tag}_theta"], rtol=0, atol=1e-16)
        if kw.get("treatment"):
            np.testing.assert_array_equal(c.theta_treatment, g[f"{tag}_theta_treatment"])
