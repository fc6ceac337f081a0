"""ctypes binding of oracle/_ref/libstm_oracle.so (the C fp64 restatement; TEST INFRASTRUCTURE)."""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "_ref", "libstm_oracle.so")
_lib = None


def build(force=False):
    """Compile the C oracle with the committed Makefile (gcc only)."""
    if force or not os.path.exists(_SO) or os.path.getmtime(_SO) < os.path.getmtime(
            os.path.join(_HERE, "stm_oracle.c")):
        subprocess.run(["make", "-C", _HERE, "-B"], check=True, capture_output=True)
    return _SO


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        dp, ip, lp = C.POINTER(C.c_double), C.POINTER(C.c_int32), C.POINTER(C.c_int64)
        L.stm_oracle_prologue.argtypes = [C.c_int, dp, dp, dp]
        L.stm_oracle_prologue.restype = C.c_int
        L.stm_oracle_estep.argtypes = [C.c_int64, C.c_int, C.c_int, C.c_int, lp, ip, dp, ip,
                                       dp, dp, dp, C.c_double, dp, dp, dp, dp, dp,
                                       dp, ip, ip, ip, ip, ip, C.c_int]
        L.stm_oracle_estep.restype = C.c_int
        L.stm_oracle_f.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp]
        L.stm_oracle_f.restype = C.c_double
        L.stm_oracle_df.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp]
        L.stm_oracle_df.restype = None
        L.stm_oracle_bfgs.argtypes = [C.c_int, C.c_int, dp, dp, dp, dp, dp, dp,
                                      C.POINTER(C.c_int), C.POINTER(C.c_int), C.POINTER(C.c_int)]
        L.stm_oracle_bfgs.restype = C.c_int
        L.stm_oracle_shortcut_check.argtypes = [C.c_int, C.POINTER(C.c_longlong)]
        L.stm_oracle_shortcut_check.restype = None
        _lib = L
    return _lib


def shortcut_check(enable):
    """Switch the oracle's self-check of the kernel's curvature certificate on / off; returns the counters gathered
    since the last call: cert_w1_fired, cert_w1_accept_after, cert_zoom_fired, cert_zoom_accept_after,
    cert_trials_skipped."""
    out = (C.c_longlong * 5)()
    lib().stm_oracle_shortcut_check(int(bool(enable)), out)
    keys = ("cert_w1_fired", "cert_w1_accept_after", "cert_zoom_fired", "cert_zoom_accept_after", "cert_trials_skipped")
    return dict(zip(keys, out))


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t)) if a is not None else None


def prologue(sigma):
    """stm.py:497-501 -> (siginv, sigmaentropy)"""
    sigma = np.ascontiguousarray(sigma, dtype=np.float64)
    K1 = sigma.shape[0]
    siginv = np.zeros((K1, K1))
    ent = C.c_double(0.0)
    rc = lib().stm_oracle_prologue(K1, _p(sigma, C.c_double), _p(siginv, C.c_double), C.byref(ent))
    if rc != 0:
        raise np.linalg.LinAlgError("sigma is not positive definite")
    return siginv, ent.value


def estep(doc_ptr, word_id, count, beta, mu, siginv, sigmaentropy, eta, aspect=None, nthreads=1):
    """Full E-step (stm.py:489-597). Returns a dict; `eta` is not modified (a copy is updated)."""
    doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
    word_id = np.ascontiguousarray(word_id, dtype=np.int32)
    count = np.ascontiguousarray(count, dtype=np.float64)
    beta = np.ascontiguousarray(beta, dtype=np.float64)
    if beta.ndim == 2:
        A, (K, V) = 1, beta.shape
    else:
        A, K, V = beta.shape
    D = doc_ptr.shape[0] - 1
    K1 = K - 1
    mu = np.ascontiguousarray(np.broadcast_to(mu, (D, K1)), dtype=np.float64)
    siginv = np.ascontiguousarray(siginv, dtype=np.float64)
    eta = np.array(eta, dtype=np.float64, order="C", copy=True).reshape(D, K1)
    asp = None if aspect is None else np.ascontiguousarray(aspect, dtype=np.int32)
    out = dict(
        eta=eta, theta=np.zeros((D, K)), beta_ss=np.zeros(beta.shape), sigma_ss=np.zeros((K1, K1)),
        doc_bound=np.zeros(D), status=np.zeros(D, np.int32), nit=np.zeros(D, np.int32),
        nfev=np.zeros(D, np.int32), njev=np.zeros(D, np.int32), repair=np.zeros(D, np.int32))
    bound = C.c_double(0.0)
    rc = lib().stm_oracle_estep(
        D, K, V, A, _p(doc_ptr, C.c_int64), _p(word_id, C.c_int32), _p(count, C.c_double),
        _p(asp, C.c_int32), _p(beta, C.c_double), _p(mu, C.c_double), _p(siginv, C.c_double),
        float(sigmaentropy), _p(eta, C.c_double), _p(out["theta"], C.c_double),
        _p(out["beta_ss"], C.c_double), _p(out["sigma_ss"], C.c_double), C.byref(bound),
        _p(out["doc_bound"], C.c_double), _p(out["status"], C.c_int32), _p(out["nit"], C.c_int32),
        _p(out["nfev"], C.c_int32), _p(out["njev"], C.c_int32), _p(out["repair"], C.c_int32),
        int(nthreads))
    if rc != 0:
        raise ValueError(f"stm_oracle_estep failed with code {rc}")
    out["bound"] = bound.value
    return out


def f(beta_doc, count, mu, siginv, eta):
    beta_doc = np.ascontiguousarray(beta_doc, dtype=np.float64)
    K, n = beta_doc.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (count, mu, siginv, eta)]
    return lib().stm_oracle_f(K, n, _p(beta_doc, C.c_double), *[_p(x, C.c_double) for x in a])


def df(beta_doc, count, mu, siginv, eta):
    beta_doc = np.ascontiguousarray(beta_doc, dtype=np.float64)
    K, n = beta_doc.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (count, mu, siginv, eta)]
    g = np.zeros(K - 1)
    lib().stm_oracle_df(K, n, _p(beta_doc, C.c_double), *[_p(x, C.c_double) for x in a], _p(g, C.c_double))
    return g


def bfgs(beta_doc, count, mu, siginv, eta0):
    """scipy.optimize.minimize(f, eta0, jac=df, method='BFGS') restated. -> dict(x, fun, status, nit, nfev, njev)"""
    beta_doc = np.ascontiguousarray(beta_doc, dtype=np.float64)
    K, n = beta_doc.shape
    a = [np.ascontiguousarray(x, dtype=np.float64) for x in (count, mu, siginv)]
    x = np.array(eta0, dtype=np.float64, copy=True)
    fun = C.c_double(0.0)
    nit, nfev, njev = C.c_int(0), C.c_int(0), C.c_int(0)
    st = lib().stm_oracle_bfgs(K, n, _p(beta_doc, C.c_double), *[_p(v, C.c_double) for v in a],
                               _p(x, C.c_double), C.byref(fun), C.byref(nit), C.byref(nfev), C.byref(njev))
    return dict(x=x, fun=fun.value, status=st, nit=nit.value, nfev=nfev.value, njev=njev.value)
