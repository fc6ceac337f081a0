"""ctypes binding of libstm_b200.so (include/stm_b200.h).  There is NO CPU fallback: a missing
library or a missing GPU raises."""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libstm_b200.so")

STM_OK = 0
STM_ERR_INVALID, STM_ERR_CUDA, STM_ERR_NOT_PD, STM_ERR_UNSUPPORTED, STM_ERR_NO_CORPUS = -1, -2, -3, -4, -5
MODEL_STM, MODEL_CTM, MODEL_STM_RIDGE, MODEL_STM_LASSO = 0, 1, 2, 3

EXPORTS = [
    "stm_create", "stm_destroy", "stm_last_error", "stm_beta_stride", "stm_launch_count", "stm_estep_kernel_ms",
    "stm_tune",
    "stm_set_corpus", "stm_stats_layout", "stm_prologue", "stm_estep", "stm_moments", "stm_mstep",
    "stm_beta_to_wordmajor", "stm_wordmajor_to_kv", "stm_estep_host", "stm_heldout", "stm_heldout64", "stm_heldout_host",
    "stm_spectral_gram", "stm_spectral_finish", "stm_sample_corpus", "stm_update_kappa",
]

_lib = None
# launch-configuration overrides applied to every new Context (tools/ A/B runs set this; empty = library defaults)
DEFAULT_TUNE = {}


class StmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"libstm_b200 error {code}: {msg}")
        self.code = code


def load():
    """Loads the shared library (built by `__graft_entry__.build()` / `make -C strutopy_b200/csrc`)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise ImportError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(strutopy_b200 has no CPU fallback)")
    L = C.CDLL(LIB_PATH)
    vp, i32, i64, dbl = C.c_void_p, C.c_int, C.c_int64, C.c_double
    L.stm_create.argtypes = [i32, i32, i32, i32, C.POINTER(vp)]
    L.stm_destroy.argtypes = [vp]
    L.stm_destroy.restype = None
    L.stm_last_error.argtypes = [vp]
    L.stm_last_error.restype = C.c_char_p
    L.stm_beta_stride.argtypes = [i32]
    L.stm_launch_count.argtypes = [vp]
    L.stm_launch_count.restype = i64
    L.stm_estep_kernel_ms.argtypes = [vp, C.POINTER(C.c_double)]
    L.stm_tune.argtypes = [vp, C.c_char_p, i32]
    L.stm_heldout.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.stm_heldout64.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, vp, vp]
    L.stm_heldout_host.argtypes = [vp, i64, vp, vp, vp, vp, vp, vp, C.POINTER(C.c_double)]
    L.stm_spectral_gram.argtypes = [vp, i32, vp, vp, vp]
    L.stm_spectral_finish.argtypes = [vp, i32, vp, vp, vp, vp, vp, vp]
    L.stm_sample_corpus.argtypes = [vp, i64, i32, vp, vp, C.c_uint64, vp, vp, vp, C.POINTER(i64), vp]
    L.stm_update_kappa.argtypes = [vp, vp, vp, dbl, i32, vp, vp, vp, vp]
    L.stm_set_corpus.argtypes = [vp, i64, vp, vp, vp, vp]
    L.stm_stats_layout.argtypes = [vp, i32, C.POINTER(i64)]
    L.stm_prologue.argtypes = [vp, vp, vp, vp, vp]
    L.stm_estep.argtypes = [vp] * 11
    L.stm_moments.argtypes = [vp, vp, vp, i32, vp, vp]
    L.stm_mstep.argtypes = [vp, vp, vp, i32, i32, dbl, vp, vp, vp, vp, vp, vp]
    L.stm_beta_to_wordmajor.argtypes = [vp, vp, vp, vp]
    L.stm_wordmajor_to_kv.argtypes = [vp, vp, vp, vp]
    L.stm_estep_host.argtypes = [vp, vp, vp, vp, dbl] + [vp] * 9
    for name in EXPORTS:
        if name not in ("stm_destroy", "stm_last_error", "stm_launch_count"):
            getattr(L, name).restype = i32
    _lib = L
    return L


def check(ctx, rc):
    if rc != STM_OK:
        msg = load().stm_last_error(ctx)
        raise StmError(rc, msg.decode() if msg else "")


def hp(a):
    """host pointer of a C-contiguous numpy array (or None)"""
    if a is None:
        return None
    assert a.flags["C_CONTIGUOUS"]
    return C.c_void_p(a.ctypes.data)


class Context:
    """Owns one stm_ctx (one GPU)."""

    def __init__(self, K, V, A=1, device=0, tune=None):
        L = load()
        self._h = C.c_void_p()
        rc = L.stm_create(int(device), int(K), int(V), int(A), C.byref(self._h))
        if rc != STM_OK:
            msg = L.stm_last_error(None)
            raise StmError(rc, msg.decode() if msg else "")
        for key, value in (tune or DEFAULT_TUNE).items():
            check(self._h, L.stm_tune(self._h, key.encode(), int(value)))
        self.K, self.V, self.A, self.device = int(K), int(V), int(A), int(device)
        self.K1 = self.K - 1
        self.TS = L.stm_beta_stride(self.K)
        self.D = 0

    @property
    def handle(self):
        return self._h

    def close(self):
        if self._h:
            load().stm_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def set_corpus(self, doc_ptr, word_id, count, aspect=None):
        doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
        word_id = np.ascontiguousarray(word_id, dtype=np.int32)
        count = np.ascontiguousarray(count, dtype=np.float32)
        asp = None if aspect is None else np.ascontiguousarray(aspect, dtype=np.int32)
        D = doc_ptr.shape[0] - 1
        check(self._h, load().stm_set_corpus(self._h, D, hp(doc_ptr), hp(word_id), hp(count), hp(asp)))
        self.D = D

    def stats_layout(self, p):
        off = (C.c_int64 * 10)()
        check(self._h, load().stm_stats_layout(self._h, int(p), off))
        return [int(x) for x in off]

    def launch_count(self):
        return int(load().stm_launch_count(self._h))

    def heldout_host(self, doc_ptr, word_id, count, theta, beta):
        """eval_heldout (heldout.py:88-97) with host arrays -> (mean, per-document per-word log-likelihood)"""
        doc_ptr = np.ascontiguousarray(doc_ptr, dtype=np.int64)
        word_id = np.ascontiguousarray(word_id, dtype=np.int32)
        count = np.ascontiguousarray(count, dtype=np.float32)
        D = doc_ptr.shape[0] - 1
        theta = np.ascontiguousarray(theta, dtype=np.float64)
        beta = np.ascontiguousarray(beta, dtype=np.float64)
        if theta.shape != (D, self.K) or beta.shape != (self.K, self.V):
            raise ValueError(f"theta must be ({D}, {self.K}) and beta ({self.K}, {self.V})")
        doc_ll = np.empty(D, dtype=np.float64)
        mean = C.c_double()
        check(self._h, load().stm_heldout_host(self._h, D, hp(doc_ptr), hp(word_id), hp(count), hp(theta), hp(beta),
                                               hp(doc_ll), C.byref(mean)))
        return float(mean.value), doc_ll

    def estep_kernel_ms(self):
        """(kernel A ms, kernel B ms) of the last E-step on this context (CUDA events on its stream)."""
        ms = (C.c_double * 2)()
        check(self._h, load().stm_estep_kernel_ms(self._h, ms))
        return float(ms[0]), float(ms[1])

    def estep_host(self, beta, mu, siginv, sigmaentropy, eta, want_docs=True):
        """E_step() with host fp64 arrays in the reference's layout (stm.py:489-597)."""
        K, K1, V, A, D = self.K, self.K1, self.V, self.A, self.D
        beta = np.ascontiguousarray(beta, dtype=np.float64).reshape(A, K, V)
        mu = np.ascontiguousarray(np.broadcast_to(mu, (D, K1)), dtype=np.float64)
        siginv = np.ascontiguousarray(siginv, dtype=np.float64).reshape(K1, K1)
        eta = np.array(eta, dtype=np.float64, order="C", copy=True).reshape(D, K1)
        out = dict(eta=eta, theta=np.empty((D, K)), beta_ss=np.empty((A, K, V)),
                   sigma_ss=np.empty((K1, K1)))
        bound = np.zeros(1)
        if want_docs:
            out.update(doc_bound=np.empty(D), status=np.empty(D, np.int32), nit=np.empty(D, np.int32),
                       repair=np.empty(D, np.int32))
        check(self._h, load().stm_estep_host(
            self._h, hp(beta), hp(mu), hp(siginv), float(sigmaentropy), hp(eta), hp(out["theta"]),
            hp(out["beta_ss"]), hp(out["sigma_ss"]), hp(bound), hp(out.get("doc_bound")),
            hp(out.get("status")), hp(out.get("nit")), hp(out.get("repair"))))
        out["bound"] = float(bound[0])
        if A == 1:
            out["beta_ss"] = out["beta_ss"][0]
        return out
