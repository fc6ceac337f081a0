"""Drop-in `STM` front for the reference class /root/reference/src/modules/stm.py:310-1259.

Same constructor keywords, methods and attributes; the inside of `E_step` / `M_step` runs on a B200
through libstm_b200.so (include/stm_b200.h).  PyTorch tensors are only the owners of device memory
(and `torch.distributed` the plumbing for the one all-reduce per EM iteration).  State lives on the
device; `beta`, `eta`, `theta`, `mu`, `sigma`, `gamma` are host views fetched on access and uploaded
on assignment, so reference-style state injection (`model.beta = ...`) keeps working.

New keyword arguments (all optional): `device`; `distributed` (shard the documents over the ranks of the
initialised torch.distributed group); `presharded` (with `distributed`: `documents`, `X`, `beta_index` already are
this rank's shard).  Their defaults reproduce the reference.  ONE default deviates from the reference on purpose:
`mnreg_column` (content model, `lda_beta=False` only).  The reference's `mnreg` regresses EVERY word on count column 1
(stm.py:825), which makes beta collapse to the unigram distribution, and it raises on SciPy >= 1.14; the default
`mnreg_column=None` regresses word v on its own column (the minimal repair, logged as a warning on first use, pinned
against the oracle in tests/test_mnreg.py::test_gpu_kappa_own_column_vs_oracle_and_fit); `mnreg_column=1` reproduces
the reference AS WRITTEN (pinned against the live reference) — DESIGN.md §10.4.

The array attributes are read-only snapshots: assign the whole attribute (`model.eta = new`) to change device state;
an in-place edit (`model.eta[i] = x`) raises instead of silently changing only the host copy.
"""
import logging
import os
import pickle
import time
from operator import itemgetter

import numpy as np

from . import _lib
from .corpus import pack_corpus, word_counts
from .parallel import allreduce_stats, shard_bounds

logger = logging.getLogger(__name__)


def _torch():
    import torch
    if not torch.cuda.is_available():
        raise RuntimeError("strutopy_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback")
    return torch


def design_matrix(X):
    """Covariates as the 2-D design matrix update_mu regresses on (stm.py:656-671): kept as is when
    exactly 0/1, otherwise one-hot encoded column by column (sorted categories, like sklearn's
    OneHotEncoder)."""
    try:
        X = X.astype("category")  # mirrors stm.py:656-659 (only succeeds for pandas objects)
    except Exception:
        pass
    cov = np.array(X)[:, None]
    if cov.ndim > 2:
        cov = np.squeeze(cov, axis=1)
    if not np.array_equal(cov, cov.astype(bool)):
        cols = []
        for j in range(cov.shape[1]):
            cats = np.unique(cov[:, j])
            cols.append((cov[:, j][:, None] == cats[None, :]).astype(np.float64))
        cov = np.concatenate(cols, axis=1)
    return np.ascontiguousarray(cov, dtype=np.float64)


class STM:
    def __init__(self, documents, dictionary, content, K, X, kappa_interactions, max_em_iter,
                 sigma_prior, convergence_threshold, lda_beta=True, beta_index=None, A=None,
                 dtype=np.float32, init_type="spectral", model_type="STM", mode="ols",
                 device=None, distributed=False, presharded=False, mnreg_column=None):
        """Keyword-compatible with stm.py:311-329.  `documents`: list of [(word_id, count), ...] (or a
        pre-packed CSR triple); `dictionary`: anything with len() and item lookup."""
        np.random.seed(123456)  # stm.py:361 (`random` there is numpy.random)
        self.dtype = np.finfo(dtype).dtype
        self.documents = documents
        self.dictionary = dictionary
        self.init = init_type
        self.model = model_type
        self.mode = mode
        self.content = content
        self.K = K
        self.A = A
        self.V = len(self.dictionary)
        self.interactions = kappa_interactions
        self.beta_index = beta_index
        self.betaindex = beta_index
        self.max_em_iter = max_em_iter
        self.max_em_its = max_em_iter
        self.sigma_prior = sigma_prior
        self.convergence_threshold = convergence_threshold
        self.N = len(self.documents) if not isinstance(documents, tuple) else len(documents[0]) - 1
        self.LDAbeta = lda_beta
        self.last_bounds = []
        self.bound = None
        self.time_processed = None

        if self.K == 0 or self.K is None:
            raise ValueError("Number of topics must be specified")  # stm.py:393-394
        if self.A == 1:
            logging.warning("no dimension for the topical content provided")
        self.mnreg_column = mnreg_column
        if not self.LDAbeta and not (self.interactions and self.A and int(self.A) >= 2):
            raise NotImplementedError(
                "lda_beta=False (STM.mnreg, stm.py:749-853) is the CONTENT model: it needs kappa_interactions=True, "
                "A >= 2 and beta_index (the reference's own A == 1 branch indexes beta_ss[1] of a K x V array)")
        if self.model not in ("STM", "CTM"):
            raise ValueError('Updating the topical prevalence parameter requires a mode. Choose from '
                             '"CTM", "Pooled" or "L1" (default).')  # stm.py:708-711

        torch = _torch()
        self._torch = torch
        # ---- distribution over GPUs: documents sharded, one all-reduce per EM iteration ---------
        self._dist = None
        self.rank, self.world = 0, 1
        if distributed:
            import torch.distributed as dist
            if not dist.is_initialized():
                raise RuntimeError("distributed=True needs an initialised torch.distributed process group")
            self._dist = dist
            self.rank, self.world = dist.get_rank(), dist.get_world_size()
        if device is None:
            device = int(os.environ.get("LOCAL_RANK", "0")) if distributed else torch.cuda.current_device()
        self.device = int(device)
        self._dev = torch.device("cuda", self.device)

        # ---- corpus: CSR, resident in HBM for the whole fit ------------------------------------
        ptr, ids, cnt = pack_corpus(documents)
        self._use_aspect = bool(self.interactions)
        nA = int(self.A) if (self._use_aspect and self.A) else 1
        self._nA = nA
        self.wcounts = word_counts(ptr, ids, cnt, self.V)  # stm.py:485-486
        aspect = None
        if self._use_aspect:
            if beta_index is None:
                raise ValueError("kappa_interactions=True needs beta_index (one content level per document)")
            aspect = np.ascontiguousarray(np.asarray(beta_index), dtype=np.int32)
        if self._dist is not None and presharded:
            lo, hi = 0, self.N
            cnt_t = torch.tensor([self.N], dtype=torch.int64, device=self._dev)
            allreduce_stats(cnt_t, self._dist)
            self._presharded_total = int(cnt_t.item())
        else:
            lo, hi = shard_bounds(ptr, self.world)[self.rank]
            self._presharded_total = None
        self._lo, self._hi = lo, hi
        self.N_local = hi - lo
        lptr = ptr[lo:hi + 1] - ptr[lo]
        lids = ids[ptr[lo]:ptr[hi]]
        lcnt = cnt[ptr[lo]:ptr[hi]]
        self._ctx = _lib.Context(self.K, self.V, nA, self.device)
        self._ctx.set_corpus(lptr, lids, lcnt, None if aspect is None else aspect[lo:hi])

        # ---- prevalence design matrix (constant over the fit) ----------------------------------
        self.X = X
        if self.model == "STM":
            self._design = design_matrix(X)
            if self._design.shape[0] != self.N:
                raise ValueError("X must have one row per document")
        else:
            self._design = np.zeros((self.N, 0))
        self._p = int(self._design.shape[1])
        if self._presharded_total is not None:
            self.N = self._presharded_total  # global document count; local rows are [0, N_local)

        # ---- device state -------------------------------------------------------------------------
        K1, Dl, TS = self.K - 1, self.N_local, self._ctx.TS
        f64 = dict(dtype=torch.float64, device=self._dev)
        self._off = self._ctx.stats_layout(self._p)
        self._d = dict(
            beta_t=torch.zeros((nA, self.V, TS), dtype=torch.float32, device=self._dev),
            # fp64 master copy of beta (word-major like beta_t): what `beta`, save_model and eval_heldout report;
            # the E-step kernels read the fp32 copy (north-star storage format)
            beta64_t=torch.zeros((nA, self.V, TS), dtype=torch.float64, device=self._dev),
            mu=torch.zeros((Dl, K1), **f64),
            eta=torch.zeros((Dl, K1), **f64),
            theta=torch.zeros((Dl, self.K), **f64),
            sigma=torch.zeros((K1, K1), **f64),
            prior=torch.zeros(K1 + 1, **f64),
            stats=torch.zeros(self._off[9], **f64),
            doc_bound=torch.zeros(max(Dl, 1), **f64),
            doc_info=torch.zeros(max(Dl, 1), dtype=torch.int32, device=self._dev),
            doc_nfev=torch.zeros(max(Dl, 1), dtype=torch.int32, device=self._dev),
            info=torch.zeros(4, dtype=torch.int32, device=self._dev),
            x=torch.from_numpy(self._design[lo:hi].copy()).to(self._dev),
            gamma_t=torch.zeros((max(self._p, 1), K1), **f64),
        )
        self._host = {}
        self._have_gamma = False
        self.init_params()

    # ------------------------------------------------------------------------------------------------
    # host views of device state
    # ------------------------------------------------------------------------------------------------
    def _stream(self):
        return self._torch.cuda.current_stream(self._dev).cuda_stream

    def _ptr(self, name):
        return self._d[name].data_ptr()

    def _invalidate(self, *names):
        for n in names:
            self._host.pop(n, None)

    def _gather_rows(self, t):
        """local [D_local, c] device tensor -> full [N, c] host array (all ranks)."""
        if self._dist is None:
            return t.cpu().numpy()
        parts = [None] * self.world
        self._dist.all_gather_object(parts, t.cpu().numpy())
        return np.concatenate(parts, axis=0)

    def _ro(self, a):
        """host views are snapshots of device state: read-only, so that an in-place edit (model.eta[i] = x) raises
        instead of silently changing only the cache — assign the whole attribute to change device state"""
        a.setflags(write=False)
        return a

    @property
    def beta(self):
        """K x V (A x K x V with a content covariate) float64 — the fp64 master copy the M-step writes"""
        if "beta" not in self._host:
            b = self._d["beta64_t"][:, :, :self.K].permute(0, 2, 1).contiguous().cpu().numpy()
            self._host["beta"] = self._ro(b if self._use_aspect else b[0])
        return self._host["beta"]

    @beta.setter
    def beta(self, value):
        torch = self._torch
        b = np.asarray(value, dtype=np.float64).reshape(self._nA, self.K, self.V)
        src = torch.from_numpy(np.ascontiguousarray(b)).to(self._dev)
        self._d["beta64_t"].zero_()
        self._d["beta64_t"][:, :, :self.K].copy_(src.permute(0, 2, 1))
        _lib.check(self._ctx.handle, _lib.load().stm_beta_to_wordmajor(
            self._ctx.handle, src.data_ptr(), self._ptr("beta_t"), self._stream()))
        torch.cuda.current_stream(self._dev).synchronize()
        self._invalidate("beta")

    @property
    def kappa(self):
        """content-covariate coefficients (stm.py:841), p x V; None before the first update (fetched on access)"""
        if "kappa" not in self._d:
            return None
        if "kappa" not in self._host:
            self._host["kappa"] = self._ro(self._d["kappa"].cpu().numpy())
        return self._host["kappa"]

    @property
    def gamma(self):
        """K-1 x p prevalence coefficients (stm.py:703); None before the first M-step, as in the reference"""
        if not self._have_gamma:
            return None
        if "gamma" not in self._host:
            self._host["gamma"] = self._ro(self._d["gamma_t"][:self._p].t().contiguous().cpu().numpy())
        return self._host["gamma"]

    def _get_rows(self, name):
        if name not in self._host:
            self._host[name] = self._ro(self._gather_rows(self._d[name]))
        return self._host[name]

    def _set_rows(self, name, value, cols):
        value = np.asarray(value, dtype=np.float64)
        if self._presharded_total is not None:
            v = np.ascontiguousarray(np.broadcast_to(value, (self.N_local, cols)))
        else:
            v = np.ascontiguousarray(np.broadcast_to(value, (self.N, cols)))[self._lo:self._hi]
        self._d[name].copy_(self._torch.from_numpy(np.array(v, dtype=np.float64, order="C", copy=True)))
        self._invalidate(name)

    eta = property(lambda s: s._get_rows("eta"), lambda s, v: s._set_rows("eta", v, s.K - 1))
    mu = property(lambda s: s._get_rows("mu"), lambda s, v: s._set_rows("mu", v, s.K - 1))
    theta = property(lambda s: s._get_rows("theta"), lambda s, v: s._set_rows("theta", v, s.K))

    @property
    def sigma(self):
        if "sigma" not in self._host:
            self._host["sigma"] = self._ro(self._d["sigma"].cpu().numpy())
        return self._host["sigma"]

    @sigma.setter
    def sigma(self, value):
        v = np.ascontiguousarray(np.asarray(value, dtype=np.float64).reshape(self.K - 1, self.K - 1))
        self._d["sigma"].copy_(self._torch.from_numpy(v))
        self._invalidate("sigma")

    # ------------------------------------------------------------------------------------------------
    # initialisation — stm.py:402-486
    # ------------------------------------------------------------------------------------------------
    def init_params(self):
        self.init_beta()
        self.init_mu()
        self.init_eta()
        self.init_sigma()
        self.init_theta()

    def init_beta(self):
        if self.init == "spectral":
            # stm.py:420-423 -> spectral_init(documents, K, V, maxV=5000): Gram statistics of the local
            # shard, one all-reduce, anchors + recovery replicated on every rank
            from .spectral import spectral_on_context
            totals = np.asarray(self.wcounts, dtype=np.float64)
            if self._presharded_total is not None:
                t = self._torch.from_numpy(totals.copy()).to(self._dev)
                allreduce_stats(t, self._dist)
                totals = t.cpu().numpy()
            width = int(np.flatnonzero(totals).max()) + 1 if np.any(totals) else 1   # create_dtm's width, stm.py:119
            b = spectral_on_context(self._ctx, self._torch, self._dev, totals[:width], maxV=5000, dist=self._dist)
        elif self.init == "random":
            # stm.py:425-429: gamma(0.1, 1) from the legacy RNG seeded in the constructor, row-normalised
            b = np.random.gamma(0.1, 1, self.V * self.K).reshape(self.K, self.V)
            rs = np.sum(b, axis=1)[:, None]
            b = np.divide(b, rs, out=np.zeros_like(b), where=rs != 0)
        else:
            raise ValueError("init_type must be 'spectral' or 'random'")
        if self._use_aspect:
            b = np.repeat(b[None, :], self._nA, axis=0)  # stm.py:430-431
        self.beta = b

    def init_mu(self):
        self._d["mu"].zero_()
        self._invalidate("mu")

    def init_eta(self):
        self._d["eta"].zero_()
        self._invalidate("eta")

    def init_sigma(self):
        s = self._d["sigma"]
        s.zero_()
        s.fill_diagonal_(20.0)  # stm.py:459-461
        self._invalidate("sigma")

    def init_theta(self):
        self._d["theta"].zero_()
        self._invalidate("theta")

    # ------------------------------------------------------------------------------------------------
    # E-step / M-step
    # ------------------------------------------------------------------------------------------------
    def _estep_device(self):
        """prologue + document loop on the device; asynchronous."""
        L, h, st = _lib.load(), self._ctx.handle, self._stream()
        _lib.check(h, L.stm_prologue(h, self._ptr("sigma"), self._ptr("prior"), self._ptr("info"), st))
        _lib.check(h, L.stm_estep(h, self._ptr("beta_t"), self._ptr("mu"), self._ptr("prior"), self._ptr("eta"),
                                  self._ptr("theta"), self._ptr("stats"), self._ptr("doc_bound"),
                                  self._ptr("doc_info"), self._ptr("doc_nfev"), st))
        self._invalidate("eta", "theta")

    def _allreduce(self, t):
        """the ONE collective of an EM iteration (and of M_step / the spectral Gram): sum over the document shards"""
        allreduce_stats(t, self._dist)

    def _reduce_and_bound(self):
        """moments, the one all-reduce of the packed statistics, and the ELBO — the ONE host synchronisation of an
        EM iteration (a Sigma that is not positive definite makes the prior, hence the bound, NaN: the Cholesky status
        is only fetched then)."""
        L, h, st = _lib.load(), self._ctx.handle, self._stream()
        _lib.check(h, L.stm_moments(h, self._ptr("eta"), self._ptr("x"), self._p, self._ptr("stats"), st))
        self._allreduce(self._d["stats"])
        bound = float(self._d["stats"][self._off[2]].item())
        if bound != bound and int(self._d["info"][0].item()) != 0:
            # the reference's except-branch calls logging.ERROR(...) (stm.py:503-506) and dies
            raise np.linalg.LinAlgError("Cholesky Decomposition failed, because Sigma is not positive definite.")
        return bound

    def _mstep_device(self):
        L, h, st = _lib.load(), self._ctx.handle, self._stream()
        if self.model != "STM":
            model = _lib.MODEL_CTM
        else:   # stm.py:673-694: any other mode string falls back to 'ols' (after a printed notice)
            if self.mode not in ["lasso", "ridge", "ols"]:
                print("Need to specify the estimation mode of prevalence covariate coefficients. Uses default 'ols'.")
            model = {"ridge": _lib.MODEL_STM_RIDGE, "lasso": _lib.MODEL_STM_LASSO}.get(self.mode, _lib.MODEL_STM)
        _lib.check(h, L.stm_mstep(h, self._ptr("stats"), self._ptr("x"), self._p, model, float(self.sigma_prior),
                                  self._ptr("gamma_t"), self._ptr("mu"), self._ptr("sigma"),
                                  self._ptr("beta_t"), self._ptr("beta64_t"), st))
        if self.model == "STM":
            self._have_gamma = True     # K1 x p, stm.py:703 — fetched on access (no host sync in the EM loop)
        if not self.LDAbeta:
            self._update_kappa_device()
        self._invalidate("mu", "sigma", "beta", "gamma")

    def _update_kappa_device(self):
        """update_beta with lda_beta=False -> mnreg (stm.py:746-853): kappa and beta from the reduced beta_ss."""
        torch = self._torch
        L, h, st = _lib.load(), self._ctx.handle, self._stream()
        if "logm" not in self._d:
            w = np.asarray(self.wcounts, dtype=np.float64)
            if self._presharded_total is not None:
                t = torch.from_numpy(w.copy()).to(self._dev)
                allreduce_stats(t, self._dist)
                w = t.cpu().numpy()
            with np.errstate(divide="ignore"):
                m = np.log(w) - np.log(np.sum(w))                      # stm.py:795-797
            self._d["logm"] = torch.from_numpy(m).to(self._dev)
            p_rows = self.K + self._nA + self._nA * self.K + 1
            self._d["kappa"] = torch.zeros((p_rows, self.V), dtype=torch.float64, device=self._dev)
        col = -1 if self.mnreg_column is None else int(self.mnreg_column)
        if col < 0 and not getattr(self, "_warned_mnreg", False):
            logger.warning("lda_beta=False: every word is regressed on its OWN count column (mnreg_column=None); the "
                           "reference as written uses column 1 for every word (stm.py:825) - pass mnreg_column=1 for that")
            self._warned_mnreg = True
        _lib.check(h, L.stm_update_kappa(h, self._ptr("stats"), self._ptr("logm"), 250.0, col, self._ptr("beta_t"),
                                         self._ptr("beta64_t"), self._ptr("kappa"), st))
        self._invalidate("kappa")

    def E_step(self):
        """stm.py:489-597 — returns (beta_ss, sigma_ss) as host arrays in the reference's layout."""
        start = time.time()
        self._estep_device()
        self.bound = self._reduce_and_bound_local()
        self.last_bounds.append(self.bound)
        torch = self._torch
        o0, o1 = self._off[0], self._off[1]
        K1 = self.K - 1
        stats = self._d["stats"]
        bss_t = stats[o0:o0 + self._nA * self.V * self._ctx.TS].view(self._nA, self.V, self._ctx.TS)
        beta_ss = bss_t[:, :, :self.K].permute(0, 2, 1).contiguous().cpu().numpy()
        sigma_ss = stats[o1:o1 + K1 * K1].view(K1, K1).cpu().numpy().copy()
        logger.info(f"Lower Bound: {self.bound}")
        logger.info(f"Completed E-Step in {np.round(time.time() - start, 3)} seconds. \n")
        del torch
        return (beta_ss if self._use_aspect else beta_ss[0]), sigma_ss

    def _reduce_and_bound_local(self):
        """E_step() semantics: the bound and statistics of ALL documents (reduced when distributed)."""
        return self._reduce_and_bound()

    def M_step(self, beta_ss, sigma_ss):
        """stm.py:622-634 — takes the (possibly caller-modified) statistics E_step returned."""
        start = time.time()
        torch = self._torch
        o0, o1 = self._off[0], self._off[1]
        K1, TS = self.K - 1, self._ctx.TS
        stats = self._d["stats"]
        b = np.asarray(beta_ss, dtype=np.float64).reshape(self._nA, self.K, self.V)
        bt = np.zeros((self._nA, self.V, TS))
        bt[:, :, :self.K] = np.transpose(b, (0, 2, 1))
        stats[o0:o0 + bt.size].copy_(torch.from_numpy(bt.reshape(-1)))
        stats[o1:o1 + K1 * K1].copy_(torch.from_numpy(np.ascontiguousarray(sigma_ss, dtype=np.float64).reshape(-1)))
        # moments of the current eta (E_step left them there; recompute in case eta was reassigned)
        L, h, st = _lib.load(), self._ctx.handle, self._stream()
        _lib.check(h, L.stm_moments(h, self._ptr("eta"), self._ptr("x"), self._p, self._ptr("stats"), st))
        self._allreduce(stats[self._off[4]:])
        stats[self._off[3]] = float(self.N)
        self._mstep_device()
        torch.cuda.current_stream(self._dev).synchronize()
        logger.info(f"Completed M-Step in {np.round(time.time() - start, 3)} seconds. \n")

    def expectation_maximization(self, saving, output_dir=None):
        """stm.py:855-880 — the whole loop stays on the device; one host sync per iteration (the ELBO)."""
        first = time.time()
        logger.info(f"Fit STM for {self.K} topics")
        for _iteration in range(100):  # hard cap, stm.py:859
            logger.info(f"E-Step iteration {_iteration}")
            self._estep_device()
            self.bound = self._reduce_and_bound()
            self.last_bounds.append(self.bound)
            logger.info(f"Lower Bound: {self.bound}")
            logger.info(f"M-Step iteration {_iteration}")
            self._mstep_device()
            if self.EM_is_converged(_iteration):
                self.time_processed = time.time() - first
                logger.info(f"model converged in iteration {_iteration} after {self.time_processed}s")
                break
            if self.max_its_reached(_iteration):
                self.time_processed = time.time() - first
                logger.info(f"maximum number of iterations ({self.max_em_its}) reached after {self.time_processed} seconds")
                break
        self._torch.cuda.current_stream(self._dev).synchronize()
        if saving:
            assert output_dir is not None
            self.save_model(output_dir)

    def EM_is_converged(self, _iteration, convergence=None):
        """stm.py:883-896"""
        if _iteration < 1:
            return False
        new, old = self.bound, self.last_bounds[-2]
        check = np.abs((new - old) / np.abs(old))
        logger.info(f"relative change: {check}")
        return bool(check < self.convergence_threshold)

    def max_its_reached(self, _iteration):
        """stm.py:898-903"""
        return _iteration == self.max_em_its - 1

    # ------------------------------------------------------------------------------------------------
    # diagnostics of the last E-step (per local document)
    # ------------------------------------------------------------------------------------------------
    def doc_diagnostics(self):
        info = self._d["doc_info"][:self.N_local].cpu().numpy()
        return dict(bound=self._d["doc_bound"][:self.N_local].cpu().numpy(), status=info & 0xF,
                    nit=(info >> 4) & 0xFFFFF, repair=(info >> 24) & 0xFF,
                    nfev=self._d["doc_nfev"][:self.N_local].cpu().numpy())

    # ------------------------------------------------------------------------------------------------
    # persistence / inspection — stm.py:1120-1259 (host, unchanged formats)
    # ------------------------------------------------------------------------------------------------
    def eval_heldout(self, heldout, return_doc_ll=False):
        """Held-out likelihood of the fitted model by document completion (heldout.py:88-97; the evaluate step of
        05_train.py:99-122) on the device-resident theta and the fp64 master beta: heldout[i] is scored with theta[i].

        Document-sharded fits: `heldout` is indexed like the documents the constructor got (global indices, or the
        local shard with presharded=True); every rank scores the held-out documents of ITS shard with its own theta
        rows, and the mean is taken over all ranks (sum and count all-reduced).  All ranks must call it.  The
        per-document values (return_doc_ll) are those of the local shard."""
        torch = self._torch
        if self._use_aspect:
            raise NotImplementedError("eval_heldout takes one K x V beta (no content covariate), heldout.py:88-97")
        ptr, ids, cnt = pack_corpus(list(heldout) if not isinstance(heldout, tuple) else heldout)
        D = ptr.shape[0] - 1
        n_index = self.N_local if self._presharded_total is not None else self.N
        if D < 1 or D > n_index:
            raise ValueError("held-out documents must be 1..N (document i is scored with theta[i])")
        if ids.size and (ids.min() < 0 or ids.max() >= self.V):
            raise IndexError("word id out of range [0, V)")
        if self._dist is not None and self._presharded_total is None:
            # global indexing: keep the documents of this rank's range [lo, hi); local row = i - lo
            lo, hi = min(self._lo, D), min(self._hi, D)
            ids, cnt = ids[ptr[lo]:ptr[hi]], cnt[ptr[lo]:ptr[hi]]
            ptr = ptr[lo:hi + 1] - ptr[lo]
            D = hi - lo
        dev = self._dev
        total = torch.zeros(2, dtype=torch.float64, device=dev)     # sum of per-document values, count
        res = np.zeros(0)
        if D >= 1:
            d_ptr = torch.from_numpy(np.ascontiguousarray(ptr)).to(dev)
            d_ids = torch.from_numpy(ids).to(dev) if ids.size else torch.zeros(1, dtype=torch.int32, device=dev)
            d_cnt = torch.from_numpy(cnt).to(dev) if cnt.size else torch.zeros(1, dtype=torch.float32, device=dev)
            out = torch.empty(D + 1, dtype=torch.float64, device=dev)
            _lib.check(self._ctx.handle, _lib.load().stm_heldout64(
                self._ctx.handle, D, d_ptr.data_ptr(), d_ids.data_ptr(), d_cnt.data_ptr(), self._ptr("theta"),
                self._ptr("beta64_t"), out.data_ptr(), out.data_ptr() + 8 * D, self._stream()))
            if self._dist is None:
                res = out.cpu().numpy()
                return (float(res[D]), res[:D]) if return_doc_ll else float(res[D])
            total[0] = out[:D].sum()
            total[1] = float(D)
            res = out[:D].cpu().numpy()
        allreduce_stats(total, self._dist)
        mean = float((total[0] / total[1]).item())
        return (mean, res) if return_doc_ll else mean

    def save_model(self, output_dir):
        """stm.py:1120-1149, unchanged file formats.  Document-sharded fits: every rank takes part in the gathers
        (theta / eta / mu are collectives), rank 0 alone writes the files; X is gathered when the ranks were given
        their own shards (presharded)."""
        beta, theta, sigma, eta, mu = self.beta, self.theta, self.sigma, self.eta, self.mu
        X = self.X
        if self._dist is not None and self._presharded_total is not None and X is not None:
            parts = [None] * self.world
            self._dist.all_gather_object(parts, np.asarray(X))
            X = np.concatenate(parts, axis=0)
        if self.rank != 0:
            return
        os.makedirs(output_dir, exist_ok=True)
        np.save(os.path.join(output_dir, "beta_hat"), beta)
        np.save(os.path.join(output_dir, "theta_hat"), theta)
        np.save(os.path.join(output_dir, "sigma_hat"), sigma)
        np.save(os.path.join(output_dir, "eta_hat"), eta)
        np.save(os.path.join(output_dir, "mu_hat"), mu)
        np.save(os.path.join(output_dir, "X"), X)
        if self.model == "STM":
            np.save(os.path.join(output_dir, "gamma_hat"), self.gamma)
        with open(os.path.join(output_dir, "lower_bound.pickle"), "wb") as f:
            pickle.dump(self.last_bounds, f)

    def ecdf(self, arr):
        """ECDF values of a 1-D array, stm.py:1257-1259"""
        import scipy.stats
        return scipy.stats.rankdata(arr, method="max") / arr.size

    def frex(self, w=0.5):
        """FREX scores, stm.py:1203-1219"""
        import scipy.special
        logbeta = np.log(self.beta)
        excl = logbeta - scipy.special.logsumexp(logbeta, axis=0)
        excl_ecdf = np.apply_along_axis(self.ecdf, 1, excl)
        freq_ecdf = np.apply_along_axis(self.ecdf, 1, logbeta)
        return 1.0 / (w / excl_ecdf + (1 - w) / freq_ecdf)

    def label_topics(self, topics, n, frexweight=0.5, print_labels=False):
        """Highest-probability and FREX words per topic, stm.py:1151-1201"""
        assert n >= 1, "n must be 1 or greater"
        topics = topics if topics else range(self.K)
        frex = self.frex(w=frexweight)
        prob_idx = np.argsort(-1 * self.beta)[:, :n]
        frex_idx = np.argsort(-1 * frex)[:, :n]
        out_prob, out_frex = [], []
        for k in topics:
            pw = [itemgetter(i)(self.dictionary) for i in prob_idx[k, :n]]
            fw = [itemgetter(i)(self.dictionary) for i in frex_idx[k, :n]]
            if print_labels:
                print(f"Topic {k}:\n \t Highest Prob: {pw}")
                print(f"Topic {k}:\n \t FREX: {fw}")
            out_prob.append(pw)
            out_frex.append(fw)
        return out_prob, out_frex

    def find_thoughts(self, topics, threshold=0, n=3):
        """Most representative documents per topic, stm.py:1221-1255"""
        assert n > 1, "Must request at least one returned document"
        n = min(n, self.N)
        theta = self.theta
        results = []
        for k in topics:
            order = np.argsort(-1 * theta[:, k])[:n]
            vals = -np.sort(-1 * theta[:, k])[:n]
            results.append(order[np.where(vals >= threshold)])
        return results[0] if len(results) == 1 else results
