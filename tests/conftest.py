import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box)")


def load_golden(name):
    with np.load(os.path.join(GOLDEN, name), allow_pickle=True) as z:
        return {k: z[k] for k in z.files}


def unpack_corpus(g, n_docs=None):
    """CSR arrays of a fixture that stores its corpus compactly (doc_len uint16, word_id uint16, count uint8:
    tests/golden/make_golden.py::pack_csr); optionally only the first n_docs documents."""
    ln = g["doc_len"].astype(np.int64)
    if n_docs is not None:
        ln = ln[:n_docs]
    ptr = np.zeros(len(ln) + 1, np.int64)
    np.cumsum(ln, out=ptr[1:])
    nnz = int(ptr[-1])
    return ptr, g["word_id"][:nnz].astype(np.int32), g["count"][:nnz].astype(np.float64)


@pytest.fixture(scope="session")
def golden():
    return load_golden


def random_init_beta(K, V):
    """The reference's random init (stm.py:361, 425-429): legacy RNG seeded 123456, gamma(0.1, 1)."""
    np.random.seed(123456)
    b = np.random.gamma(0.1, 1, V * K).reshape(K, V)
    rs = np.sum(b, axis=1)[:, None]
    return np.divide(b, rs, out=np.zeros_like(b), where=rs != 0)


def synthetic_corpus(D, V, K, n_words=150, seed=12345, p=1):
    """Vectorised generator with the reference DGP's distributions (generate_docs.py:180-316):
    beta ~ Dir(0.05), X ~ U{0,1}, eta ~ N(X gamma', 0.001 I), theta = softmax([eta, 0]),
    doc ~ Multinomial(n_words, theta beta).  Returns CSR + X + the generating beta."""
    rng = np.random.default_rng(seed)
    beta = rng.dirichlet(np.full(V, 0.05), K)
    X = rng.integers(0, 2, size=(D, p)).astype(np.float64)
    gamma = rng.normal(0.0, 1.0, size=(K - 1, p))
    eta = X @ gamma.T + rng.normal(0, np.sqrt(0.001), size=(D, K - 1))
    full = np.concatenate([eta, np.zeros((D, 1))], axis=1)
    theta = np.exp(full - full.max(1, keepdims=True))
    theta /= theta.sum(1, keepdims=True)
    ptr, ids, cnt = [0], [], []
    for d in range(D):
        w = rng.multinomial(n_words, theta[d] @ beta)
        nz = np.nonzero(w)[0]
        ids.append(nz.astype(np.int32))
        cnt.append(w[nz].astype(np.float64))
        ptr.append(ptr[-1] + len(nz))
    return (np.array(ptr, np.int64), np.concatenate(ids), np.concatenate(cnt), X, beta)
