"""Content-covariate update of beta, `STM.mnreg` (reference stm.py:746-853).

The reference method only runs with `csr_matrix.A` restored (removed in SciPy 1.14) and, as written,
regresses every word on count column 1 (stm.py:825).  tests/golden/mnreg.npz holds the LIVE reference's
kappa / beta under that shim.
CPU: the oracle in "as written" mode (column=1) EQUALS the live reference; sklearn's lbfgs answer and the
Newton minimiser of the same objective agree to 1e-7 in the repaired mode (each word on its own column).
GPU (`-m gpu`): stm_update_kappa through the STM front, `mnreg_column=1` against the live reference and the
default (own column) against the oracle.  Tolerances: kappa 1e-7 abs vs sklearn's lbfgs (its own gradient
tolerance is 1e-5 under a 250-strongly-convex objective), 1e-10 vs the Newton oracle; beta 5e-7 rel (fp32
storage of beta on the device).
"""
import numpy as np
import pytest

from conftest import load_golden
from oracle import mnreg_numpy as mn


def test_oracle_as_written_equals_live_reference():
    g = load_golden("mnreg.npz")
    beta, kappa = mn.mnreg(g["beta_ss"], g["wcounts"], column=1)
    np.testing.assert_array_equal(kappa, g["kappa"])
    np.testing.assert_array_equal(beta, g["beta"])
    K, A = int(g["K"]), int(g["A"])
    assert kappa.shape == (K + A + A * K + 1, int(g["V"]))
    assert not np.any(kappa[K])                                  # the reference's empty covariate column
    np.testing.assert_allclose(beta.sum(axis=2), 1.0, rtol=1e-12)


def test_oracle_newton_is_the_sklearn_minimiser():
    g = load_golden("mnreg.npz")
    b1, k1 = mn.mnreg(g["beta_ss"][:, :, :60], g["wcounts"][:60])
    b2, k2 = mn.mnreg(g["beta_ss"][:, :, :60], g["wcounts"][:60], solver="newton")
    assert np.abs(k1 - k2).max() <= 1e-7
    np.testing.assert_allclose(b1, b2, rtol=1e-6)
    assert np.abs(k2).max() > 1e-3                               # not the trivial solution


def _front(g, column):
    from strutopy_b200 import STM
    c = load_golden("estep_content.npz")
    K, V, A = int(g["K"]), int(g["V"]), int(g["A"])
    m = STM((c["doc_ptr"], c["word_id"], c["count"]), range(V), True, K, c["X"], True, 2, 0, 1e-5,
            init_type="random", model_type="STM", A=A, beta_index=c["aspect"], lda_beta=False, mnreg_column=column)
    np.testing.assert_allclose(m.wcounts, g["wcounts"])
    return m


@pytest.mark.gpu
def test_gpu_kappa_as_written_vs_live_reference():
    g = load_golden("mnreg.npz")
    m = _front(g, 1)
    m.M_step(g["beta_ss"], np.eye(int(g["K"]) - 1))
    assert np.abs(m.kappa - g["kappa"]).max() <= 1e-7
    np.testing.assert_allclose(m.beta, g["beta"], rtol=5e-7, atol=1e-30)


@pytest.mark.gpu
def test_gpu_kappa_own_column_vs_oracle_and_fit():
    g = load_golden("mnreg.npz")
    m = _front(g, None)
    m.M_step(g["beta_ss"], np.eye(int(g["K"]) - 1))
    ref_beta, ref_kappa = mn.mnreg(g["beta_ss"], g["wcounts"], solver="newton")
    assert np.abs(m.kappa - ref_kappa).max() <= 1e-10
    np.testing.assert_allclose(m.beta, ref_beta, rtol=5e-7, atol=1e-30)
    sk_beta, sk_kappa = mn.mnreg(g["beta_ss"], g["wcounts"])
    assert np.abs(m.kappa - sk_kappa).max() <= 1e-7
    # the content model fits end to end: E-step on the aspect-indexed beta, M-step through mnreg
    m2 = _front(g, None)
    m2.expectation_maximization(saving=False)
    assert len(m2.last_bounds) == 2 and np.all(np.isfinite(m2.last_bounds))
    np.testing.assert_allclose(m2.beta.sum(axis=2), 1.0, rtol=1e-5)


@pytest.mark.gpu
def test_lda_beta_false_needs_the_content_model():
    from strutopy_b200 import STM
    c = load_golden("estep_content.npz")
    with pytest.raises(NotImplementedError):
        STM((c["doc_ptr"], c["word_id"], c["count"]), range(int(c["V"])), False, int(c["K"]), c["X"], False, 2, 0,
            1e-5, init_type="random", lda_beta=False)
