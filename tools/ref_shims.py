"""Import shims that let the *unmodified* reference (``/root/reference/src/modules``) import in the
authoring container, where gensim / qpsolvers / matplotlib are absent.

This module is TOOLING for generating golden fixtures (tests/golden/make_golden.py).  It is never
imported by the product, the tests, smoke() or bench.py: ``/root/reference`` does not exist on the
GPU box.
"""
import sys
import types

import numpy as np

REFERENCE_SRC = "/root/reference/src"


class _Dictionary(dict):
    """Stand-in for gensim.corpora.Dictionary: STM only needs len() and item lookup (stm.py:375,1193)."""

    @classmethod
    def from_corpus(cls, corpus):
        max_id = -1
        for doc in corpus:
            for wid, _ in doc:
                max_id = max(max_id, int(wid))
        return cls({i: str(i) for i in range(max_id + 1)})


def _nnls_solve_qp(P, q, G=None, h=None, A=None, b=None, lb=None, ub=None, solver=None, **kw):
    """min 0.5 x'Px + q'x s.t. x <= 0 — exactly what recover_l2 (stm.py:245-285) asks quadprog for
    (G = I, h = 0, no equality constraint).  With w = -x and P = R'R (Cholesky) this is the NNLS
    problem min ||R w - R^-T q||, w >= 0: strictly convex, so the minimiser is unique and equals
    quadprog's up to solver rounding."""
    from scipy.linalg import cholesky, solve_triangular
    from scipy.optimize import nnls

    n = P.shape[0]
    assert A is None and b is None and lb is None and ub is None
    assert np.array_equal(G, np.eye(n)) and not np.any(h)
    R = cholesky(P, lower=False)
    w, _ = nnls(R, solve_triangular(R, q, trans="T", lower=False), maxiter=30 * n)
    return -w


def install():
    if REFERENCE_SRC not in sys.path:
        sys.path.insert(0, REFERENCE_SRC)
    # STM.mnreg (stm.py:825) uses csr_matrix.A, removed in SciPy 1.14: give the property back so that the
    # unmodified method runs (tests/golden/mnreg.npz)
    import scipy.sparse
    if not hasattr(scipy.sparse.csr_matrix, "A"):
        scipy.sparse.csr_matrix.A = property(lambda self: self.toarray())
    if "qpsolvers" not in sys.modules:
        m = types.ModuleType("qpsolvers")
        m.solve_qp = _nnls_solve_qp
        sys.modules["qpsolvers"] = m
    if "matplotlib" not in sys.modules:
        mpl = types.ModuleType("matplotlib")
        plt = types.ModuleType("matplotlib.pyplot")
        mpl.pyplot = plt
        sys.modules["matplotlib"] = mpl
        sys.modules["matplotlib.pyplot"] = plt
    if "gensim" not in sys.modules:
        g = types.ModuleType("gensim")
        gu = types.ModuleType("gensim.utils")
        gc = types.ModuleType("gensim.corpora")
        gd = types.ModuleType("gensim.corpora.dictionary")
        gd.Dictionary = _Dictionary
        gc.Dictionary = _Dictionary
        gc.dictionary = gd
        g.utils = gu
        g.corpora = gc
        sys.modules.update({"gensim": g, "gensim.utils": gu, "gensim.corpora": gc,
                            "gensim.corpora.dictionary": gd})


def load_reference():
    """Returns (stm_module, generate_docs_module) of the live reference."""
    install()
    import modules.generate_docs as gd
    import modules.stm as stm
    return stm, gd
