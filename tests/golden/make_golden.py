"""Generates the golden fixtures under tests/golden/ by running the LIVE, UNMODIFIED reference
(/root/reference/src/modules/stm.py, imported through tools/ref_shims.py).

Run in the authoring container only (the reference does not exist on the GPU box):

    python tests/golden/make_golden.py [name ...]

Fixtures (all inputs are stored next to the reference's outputs so tests need nothing else):
  kat_small.npz        SURVEY Appendix B: K=3, V=6, 2 docs, one E-step + one M-step (STM/ols)
  estep_K5.npz         state-injected E-steps, D=200 V~370 K=5 (config 1 shape), EM iterations 0 and 2
  estep_K20.npz        D=96  V~1800 K=20, EM iterations 0 and 1
  estep_K50.npz        D=64  V~1900 K=50, EM iterations 0 and 1 (PD-repair branch hot in iteration 0)
  estep_content.npz    A=2 aspects (content=True, kappa_interactions=True), K=8, one E-step + M-step
  em_c1.npz            config 1 (D=200 V=500 K=5, 1 covariate) full EM to convergence: ELBO trace + final state
  em_c2.npz            config 2 (D=10k V=5k K=20, 2 covariates): the reference-generated corpus + the live reference's
                       20-iteration ELBO trace on its first 300 documents
  em_k50.npz           K=50, D=2000 corpus, the live reference's spectral beta0 (fp32) + its 25-iteration ELBO trace on
                       the first 200 documents with that beta0 injected
  em_toy_ctm.npz       the reference's own tests/test_integration.py toy pipeline (K=3, CTM, 2 iterations)
  wiki_corpus.npz      the reference's shipped wiki BoW corpus + X + its shipped iteration-0 ELBOs (K=50, 70)
  mstep_modes.npz      update_mu / update_sigma of the live reference in its regularised modes (stm.py:678-688:
                       sklearn Lasso(alpha=1), Ridge(alpha=0.1)) and 'ols', on injected eta with strong
                       covariate effects (so that the Lasso coefficients are not all zero); binary and
                       one-hot-encoded (3-level) designs
  mnreg.npz            STM.mnreg (stm.py:749-853) of the live reference (csr_matrix.A restored by the shim) on the
                       content fixture's beta_ss: kappa and beta AS WRITTEN (every word regressed on column 1)
  corpus_params.npz    CorpusCreation's parameter draws (beta, gamma, metadata, eta, theta) of the live reference after
                       np.random.seed(12345), STM and LDA(+treatment) data-generating processes
  spectral.npz         spectral_init (stm.py:30-84) of the live reference, `solve_qp` shimmed by exact NNLS
                       (tools/ref_shims.py): two synthetic corpora (vocabulary truncated by maxV / not
                       truncated) and the shipped wiki corpus at K=20 (anchors + every 8th kept column)

beta in the state-injected fixtures is rounded to fp32-representable values BEFORE the reference
runs, so the fp32-beta CUDA path sees bit-identical inputs.
"""
import os
import pickle
import sys
import warnings

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.join(HERE, "..", "..", "tools"))
import ref_shims  # noqa: E402

warnings.filterwarnings("ignore")
stm_mod, gd_mod = ref_shims.load_reference()
REF_ART = "/root/reference/src/artifacts"


def to_csr(docs):
    ptr, ids, cnt = [0], [], []
    for doc in docs:
        for w, c in doc:
            ids.append(int(w))
            cnt.append(float(c))
        ptr.append(len(ids))
    return np.array(ptr, np.int64), np.array(ids, np.int32), np.array(cnt, np.float64)


def instrumented_estep(model):
    """Run the reference E_step while recording per-document bound and BFGS diagnostics through
    wrappers around its own methods (the reference code itself is untouched)."""
    rec = dict(bound=[], status=[], nit=[], nfev=[], njev=[])
    orig_lb, orig_opt = model.lower_bound, model.optimize_eta

    def lb(*a, **k):
        v = orig_lb(*a, **k)
        rec["bound"].append(float(v))
        return v

    def opt(*a, **k):
        r = orig_opt(*a, **k)
        rec["status"].append(r.status)
        rec["nit"].append(r.nit)
        rec["nfev"].append(r.nfev)
        rec["njev"].append(r.njev)
        return r

    model.lower_bound, model.optimize_eta = lb, opt
    try:
        beta_ss, sigma_ss = model.E_step()
    finally:
        model.lower_bound, model.optimize_eta = orig_lb, orig_opt
    return beta_ss, sigma_ss, {k: np.array(v) for k, v in rec.items()}


def snapshot_estep(model, prefix, out, round_beta=True, keep_m_beta=True):
    """Records inputs, runs the reference E-step + M-step, records outputs under `prefix`."""
    if round_beta:
        model.beta = np.asarray(model.beta, dtype=np.float32).astype(np.float64)
    # fp32-representable by construction when round_beta: store compactly
    out[prefix + "beta"] = np.array(model.beta, dtype=np.float32 if round_beta else np.float64)
    out[prefix + "mu"] = np.array(model.mu)
    out[prefix + "sigma"] = np.array(model.sigma)
    out[prefix + "eta0"] = np.array(model.eta)
    beta_ss, sigma_ss, rec = instrumented_estep(model)
    out[prefix + "siginv"] = np.array(model.siginv)
    out[prefix + "sigmaentropy"] = np.float64(model.sigmaentropy)
    out[prefix + "eta"] = np.array(model.eta)
    out[prefix + "theta"] = np.array(model.theta)
    out[prefix + "bound"] = np.float64(model.bound)
    out[prefix + "beta_ss"] = np.array(beta_ss)
    out[prefix + "sigma_ss"] = np.array(sigma_ss)
    out[prefix + "doc_bound"] = rec["bound"]
    out[prefix + "status"] = rec["status"].astype(np.int32)
    out[prefix + "nit"] = rec["nit"].astype(np.int32)
    model.M_step(beta_ss, sigma_ss)
    if keep_m_beta:
        out[prefix + "m_beta"] = np.array(model.beta)
    out[prefix + "m_mu"] = np.array(model.mu)
    out[prefix + "m_sigma"] = np.array(model.sigma)
    if getattr(model, "gamma", None) is not None:
        out[prefix + "m_gamma"] = np.array(model.gamma)


def synthetic(K, V, D, n_words, seed, level=1):
    np.random.seed(seed)
    corpus = gd_mod.CorpusCreation(n_topics=K, n_docs=D, n_words=n_words, V=V, level=level, dgp="STM")
    corpus.generate_documents(remove_terms=False)
    return corpus


def make_model(docs, dictionary, K, X, iters, model_type="STM", content=False, interactions=False,
               A=None, beta_index=None, thr=1e-5):
    return stm_mod.STM(documents=docs, dictionary=dictionary, content=content, K=K, X=X,
                       kappa_interactions=interactions, max_em_iter=iters, sigma_prior=0,
                       convergence_threshold=thr, init_type="random", model_type=model_type,
                       mode="ols", A=A, beta_index=beta_index)


def save(name, out):
    path = os.path.join(HERE, name)
    np.savez_compressed(path, **out)
    print(f"wrote {path} ({os.path.getsize(path) / 1024:.0f} KiB)")


# ------------------------------------------------------------------------------------------------

def kat_small():
    docs = [[(0, 3), (2, 1), (3, 2), (5, 4)], [(1, 2), (2, 2), (4, 1)]]
    X = np.array([[0], [1]])
    m = make_model(docs, {i: str(i) for i in range(6)}, 3, X, 2)
    B = np.array([[6, 1, 1, 1, 2, 1], [1, 5, 2, 1, 1, 2], [1, 1, 1, 4, 1, 4]], float)
    m.beta = B / B.sum(1, keepdims=True)
    m.mu = np.array([[0.2, -0.1], [0.0, 0.3]])
    m.sigma = np.array([[1.5, 0.4], [0.4, 0.8]])
    m.eta = np.array([[0.1, -0.2], [0.0, 0.0]])
    out = {}
    out["doc_ptr"], out["word_id"], out["count"] = to_csr(docs)
    out["X"] = X
    out["K"], out["V"] = np.int64(3), np.int64(6)
    snapshot_estep(m, "it0_", out, round_beta=False)
    save("kat_small.npz", out)


def estep_fixture(name, K, V, D, seed, iters_to_record, n_iters, keep_m_beta=True):
    corpus = synthetic(K, V, D, 150, seed)
    m = make_model(corpus.documents, corpus.dictionary, K, corpus.metadata, n_iters)
    out = {}
    out["doc_ptr"], out["word_id"], out["count"] = to_csr(corpus.documents)
    out["X"] = np.array(corpus.metadata)
    out["K"], out["V"] = np.int64(K), np.int64(m.V)
    for it in range(n_iters):
        if it in iters_to_record:
            snapshot_estep(m, f"it{it}_", out, keep_m_beta=keep_m_beta)
        else:
            bss, sss = m.E_step()
            m.M_step(bss, sss)
    out["recorded"] = np.array(sorted(iters_to_record), np.int64)
    save(name, out)


def estep_content():
    K, V, D, A = 8, 400, 48, 2
    corpus = synthetic(K, V, D, 120, 7)
    aspect = (np.arange(D) % A).astype(np.int32)
    m = make_model(corpus.documents, corpus.dictionary, K, corpus.metadata, 2, content=True,
                   interactions=True, A=A, beta_index=aspect)
    # give the two aspects different betas so the aspect gather matters
    rng = np.random.default_rng(3)
    b = m.beta * rng.uniform(0.5, 1.5, size=m.beta.shape)
    m.beta = b / b.sum(axis=2, keepdims=True)
    out = {}
    out["doc_ptr"], out["word_id"], out["count"] = to_csr(corpus.documents)
    out["X"] = np.array(corpus.metadata)
    out["aspect"] = aspect
    out["K"], out["V"], out["A"] = np.int64(K), np.int64(m.V), np.int64(A)
    snapshot_estep(m, "it0_", out)
    save("estep_content.npz", out)


def em_c1():
    """BASELINE.json configs[0]: D=200 V=500 K=5, 1 prevalence covariate, EM to convergence."""
    K, V, D = 5, 500, 200
    corpus = synthetic(K, V, D, 150, 12345)
    m = make_model(corpus.documents, corpus.dictionary, K, corpus.metadata, 100)
    out = {}
    out["doc_ptr"], out["word_id"], out["count"] = to_csr(corpus.documents)
    out["X"] = np.array(corpus.metadata)
    out["K"], out["V"] = np.int64(K), np.int64(m.V)
    out["beta0"] = np.array(m.beta)
    m.expectation_maximization(saving=False)
    out["bounds"] = np.array(m.last_bounds)
    for k in ("beta", "theta", "eta", "mu", "sigma", "gamma"):
        out["final_" + k] = np.array(getattr(m, k))
    print("em_c1: iterations", len(m.last_bounds), "final bound", m.last_bounds[-1])
    save("em_c1.npz", out)


def em_toy_ctm():
    """tests/test_integration.py:14-68 of the reference, verbatim parameters."""
    np.random.seed(42)
    K, V, N, n_words, level = 3, 200, 50, 50, 1
    gamma = np.random.multivariate_normal(np.random.standard_normal(level),
                                          np.diag(np.full(level, 0.001)), K - 1)
    corpus = gd_mod.CorpusCreation(n_topics=K, n_docs=N, n_words=n_words, V=V, level=level,
                                   dgp="STM", gamma=gamma)
    corpus.generate_documents(remove_terms=True)
    corpus.split_corpus(proportion=0.8)
    docs = corpus.train_docs
    np.random.seed(42)
    m = make_model(docs, corpus.dictionary, K, corpus.metadata[:len(docs)], 2, model_type="CTM")
    out = {}
    out["doc_ptr"], out["word_id"], out["count"] = to_csr(docs)
    out["X"] = np.array(corpus.metadata[:len(docs)])
    out["K"], out["V"] = np.int64(K), np.int64(m.V)
    out["beta0"] = np.array(m.beta)
    m.expectation_maximization(saving=False)
    out["bounds"] = np.array(m.last_bounds)
    for k in ("beta", "theta", "eta", "mu", "sigma"):
        out["final_" + k] = np.array(getattr(m, k))
    from modules.heldout import eval_heldout
    out["heldout_ll"] = np.float64(eval_heldout(corpus.test_2_docs, m.theta, m.beta))
    h_ptr, h_ids, h_cnt = to_csr(corpus.test_2_docs)
    out["heldout_doc_ptr"], out["heldout_word_id"], out["heldout_count"] = h_ptr, h_ids, h_cnt
    save("em_toy_ctm.npz", out)


def wiki_corpus():
    """The reference's shipped corpus + its shipped known-answer ELBOs."""
    import scipy.io
    mm = scipy.io.mmread(os.path.join(REF_ART, "wiki_data", "BoW_corpus.mm")).tocsr()
    mm.sort_indices()
    out = dict(doc_ptr=mm.indptr.astype(np.int64), word_id=mm.indices.astype(np.int32),
               count=mm.data.astype(np.int16), V=np.int64(mm.shape[1]))
    assert np.array_equal(out["count"].astype(np.float64), mm.data)
    for K in (50, 70):
        d = os.path.join(REF_ART, "reference_model", str(K))
        out[f"X_{K}"] = np.load(os.path.join(d, "X.npy"), allow_pickle=True)
        with open(os.path.join(d, "lower_bound.pickle"), "rb") as fh:
            out[f"shipped_bounds_{K}"] = np.array(pickle.load(fh), dtype=np.float64)
    save("wiki_corpus.npz", out)


def mstep_modes():
    """M1 in every `mode` (stm.py:673-706) on injected eta.  eta = X Gamma' + noise with |Gamma| up to 9:
    the Lasso threshold alpha * N is crossed for some (topic, covariate) pairs and not for others."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from conftest import synthetic_corpus
    out = {}
    D, V, K = 240, 300, 7
    ptr, ids, cnt, _, _ = synthetic_corpus(D, V, K, n_words=40, seed=21)
    docs = [[(int(ids[j]), int(cnt[j])) for j in range(ptr[d], ptr[d + 1])] for d in range(D)]
    out["doc_ptr"], out["word_id"], out["count"] = ptr, ids, cnt.astype(np.int16)
    out["K"], out["V"] = np.int64(K), np.int64(V)
    rng = np.random.default_rng(77)
    designs = dict(bin=rng.integers(0, 2, size=(D, 2)).astype(np.float64),   # kept as is (exactly 0/1)
                   cat=rng.integers(0, 3, size=(D, 1)).astype(np.float64))   # one-hot encoded -> 3 columns
    sigma_ss = rng.normal(size=(K - 1, K - 1))
    sigma_ss = sigma_ss @ sigma_ss.T * D
    out["sigma_ss"] = sigma_ss
    for dn, X in designs.items():
        m = make_model(docs, {i: str(i) for i in range(V)}, K, X, 2)
        cov = np.array(X)[:, None]
        cov = np.squeeze(cov, axis=1)
        if not np.array_equal(cov, cov.astype(bool)):
            cov = np.concatenate([(cov[:, [j]] == np.unique(cov[:, j])[None, :]).astype(float) for j in range(cov.shape[1])], axis=1)
        Gam = rng.choice([-9.0, -5.0, -0.7, 0.0, 0.4, 4.5, 8.0], size=(K - 1, cov.shape[1]))
        eta = cov @ Gam.T + rng.normal(0, 0.8, size=(D, K - 1))
        out[dn + "_X"], out[dn + "_eta"] = X, eta
        for mode in ("ols", "ridge", "lasso"):
            m.mode = mode
            m.eta = eta.copy()
            m.update_mu()
            m.update_sigma(sigma_ss, 0.3)
            out[f"{dn}_{mode}_gamma"], out[f"{dn}_{mode}_mu"], out[f"{dn}_{mode}_sigma"] = (
                np.array(m.gamma), np.array(m.mu), np.array(m.sigma))
            print(dn, mode, "nonzero gamma:", int(np.count_nonzero(m.gamma)), "of", m.gamma.size)
    save("mstep_modes.npz", out)


def mnreg():
    g = dict(np.load(os.path.join(HERE, "estep_content.npz")))
    K, V, A = int(g["K"]), int(g["V"]), int(g["A"])
    D = len(g["doc_ptr"]) - 1
    docs = [[(int(g["word_id"][j]), int(g["count"][j])) for j in range(g["doc_ptr"][d], g["doc_ptr"][d + 1])]
            for d in range(D)]
    m = make_model(docs, {i: str(i) for i in range(V)}, K, g["X"], 2, content=True, interactions=True, A=A,
                   beta_index=g["aspect"])
    m.LDAbeta = False
    out = dict(K=np.int64(K), V=np.int64(V), A=np.int64(A), beta_ss=g["it0_beta_ss"], wcounts=np.array(m.wcounts))
    m.update_beta(out["beta_ss"])                          # -> mnreg (stm.py:746-747)
    out["kappa"] = np.array(m.kappa)
    out["beta"] = np.stack(m.beta, axis=0)
    save("mnreg.npz", out)


def corpus_params():
    out = {}
    cfgs = dict(stm=dict(dgp="STM", level=2), lda=dict(dgp="LDA", level=1),
                ldat=dict(dgp="LDA", level=1, treatment=True, alpha_treatment="auto-linear", alpha="asymmetric"))
    for tag, kw in cfgs.items():
        np.random.seed(12345)
        c = gd_mod.CorpusCreation(n_topics=5, n_docs=40, n_words=30, V=60, **kw)
        for k in ("alpha", "beta", "gamma", "metadata", "eta", "theta"):
            out[f"{tag}_{k}"] = np.asarray(getattr(c, k))
        if kw.get("treatment"):
            out[f"{tag}_theta_treatment"] = np.asarray(c.theta_treatment)
    save("corpus_params.npz", out)


def spectral():
    """spectral_init of the live reference.  Case t: V=900 > maxV=500 (the `keep` cut is exercised);
    case f: every word kept; case w: the shipped wiki corpus, K=20, maxV=5000 as STM.init_beta calls it."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from conftest import synthetic_corpus
    out = {}
    for tag, (D, V, K, maxV, n_words, seed) in dict(t=(400, 900, 8, 500, 60, 7), f=(1500, 120, 6, 5000, 80, 11)).items():
        ptr, ids, cnt, _, _ = synthetic_corpus(D, V, K, n_words=n_words, seed=seed)
        if maxV >= V:  # no truncation: every word must occur (stm.py:152-154 asserts on empty rows) -> compact ids
            seen = np.unique(ids)
            ids = np.searchsorted(seen, ids).astype(np.int32)
            V = len(seen)
        docs = [[(int(ids[j]), int(cnt[j])) for j in range(ptr[d], ptr[d + 1])] for d in range(D)]
        dtm = stm_mod.create_dtm(docs)
        wprob = np.array(np.sum(dtm, axis=0) / np.sum(dtm)).flatten()
        keep = np.argsort(-1 * wprob)[:maxV]
        Q = stm_mod.gram(dtm[:, keep])
        out[tag + "_Q"] = Q.toarray()
        out[tag + "_anchor"] = np.intp(stm_mod.fastAnchor(Q, K, verbose=False))
        out[tag + "_beta"] = stm_mod.spectral_init(docs, K, V, maxV=maxV, verbose=False)
        out[tag + "_keep"] = keep
        out[tag + "_cfg"] = np.array([D, V, K, maxV, n_words, seed])
        out[tag + "_doc_ptr"], out[tag + "_word_id"], out[tag + "_count"] = ptr, ids, cnt.astype(np.int16)
    import scipy.io
    mm = scipy.io.mmread(os.path.join(REF_ART, "wiki_data", "BoW_corpus.mm")).tocsr()
    mm.sort_indices()
    docs = [[(int(mm.indices[j]), int(mm.data[j])) for j in range(mm.indptr[d], mm.indptr[d + 1])]
            for d in range(mm.shape[0])]
    K, V = 20, mm.shape[1]
    dtm = stm_mod.create_dtm(docs)
    wprob = np.array(np.sum(dtm, axis=0) / np.sum(dtm)).flatten()
    keep = np.argsort(-1 * wprob)[:5000]
    out["w_anchor"] = np.intp(stm_mod.fastAnchor(stm_mod.gram(dtm[:, keep]), K, verbose=False))
    beta = stm_mod.spectral_init(docs, K, V, maxV=5000, verbose=False)
    out["w_keep"] = keep
    out["w_cols"] = keep[::8]
    out["w_beta_cols"] = beta[:, keep[::8]]
    out["w_beta_rowsum"] = beta.sum(axis=1)
    out["w_beta_dropped"] = np.float64(beta[0, np.setdiff1d(np.arange(V), keep)[0]])
    out["w_cfg"] = np.array([mm.shape[0], V, K, 5000])
    save("spectral.npz", out)


def pack_csr(ptr, ids, cnt):
    """compact storage of a corpus: lengths uint16, ids uint16, counts uint8 (checked)"""
    ln = np.diff(ptr)
    assert ln.max() < 65536 and ids.max() < 65536 and cnt.max() < 256 and np.all(cnt == np.round(cnt))
    return ln.astype(np.uint16), ids.astype(np.uint16), cnt.astype(np.uint8)


def em_c2():
    """BASELINE.json configs[1]: D=10k, V=5k, K=20, 2 prevalence covariates, 20 EM iterations.  The corpus is drawn
    by the live reference's CorpusCreation exactly as src/04_create_synthetic_corpora.py does (np.random.seed(12345),
    150 words per document, remove_terms=False) and stored whole; the live reference itself is run on the first
    300 documents (a full fit would take ~2 h on one core) — the anchor for the C oracle's and the CUDA path's
    20-iteration traces; the full-size traces are compared GPU vs C oracle at test time."""
    K, V, D, CUT = 20, 5000, 10000, 300
    np.random.seed(12345)
    corpus = gd_mod.CorpusCreation(n_topics=K, n_docs=D, n_words=150, V=V, level=2, dgp="STM")
    corpus.generate_documents(remove_terms=False)
    ptr, ids, cnt = to_csr(corpus.documents)
    out = {}
    out["doc_len"], out["word_id"], out["count"] = pack_csr(ptr, ids, cnt)
    out["X"] = np.array(corpus.metadata, dtype=np.float64)
    Vd = len(corpus.dictionary)
    out["K"], out["V"], out["cut"] = np.int64(K), np.int64(Vd), np.int64(CUT)
    m = make_model(corpus.documents[:CUT], corpus.dictionary, K, np.array(corpus.metadata)[:CUT], 20)
    out["cut_beta0"] = np.array(m.beta)
    m.expectation_maximization(saving=False)
    out["cut_bounds"] = np.array(m.last_bounds)
    for k in ("gamma", "sigma"):
        out["cut_final_" + k] = np.array(getattr(m, k))
    out["cut_final_theta"] = np.array(m.theta)
    print("em_c2: V", Vd, "nnz", len(ids), "cut iterations", len(m.last_bounds), "final bound", m.last_bounds[-1])
    save("em_c2.npz", out)


def em_k50():
    """K=50 from a spectral initialisation, 25 EM iterations (BASELINE config 3's regime at a size the oracle runs in
    seconds): D=2000, V=2000 corpus from tests/conftest.synthetic_corpus; beta0 = the live reference's spectral_init
    (solve_qp shimmed by exact NNLS) on the whole corpus, rounded to fp32; the live reference's EM trace on the
    first 200 documents with that beta0 injected is the anchor."""
    sys.path.insert(0, os.path.join(HERE, ".."))
    from conftest import synthetic_corpus
    K, V, D, CUT = 50, 2000, 2000, 200
    ptr, ids, cnt, X, _ = synthetic_corpus(D, V, K, n_words=150, seed=4242)
    seen = np.unique(ids)                       # every word must occur (stm.py:152-154) -> compact ids
    ids = np.searchsorted(seen, ids).astype(np.int32)
    V = len(seen)
    docs = [[(int(ids[j]), int(cnt[j])) for j in range(ptr[d], ptr[d + 1])] for d in range(D)]
    beta0 = stm_mod.spectral_init(docs, K, V, maxV=5000, verbose=False)
    beta0 = np.asarray(beta0, dtype=np.float32)
    out = {}
    out["doc_len"], out["word_id"], out["count"] = pack_csr(ptr, ids, cnt)
    out["X"] = X
    out["K"], out["V"], out["cut"] = np.int64(K), np.int64(V), np.int64(CUT)
    out["beta0"] = beta0
    dictionary = {i: str(i) for i in range(V)}
    m = make_model(docs[:CUT], dictionary, K, X[:CUT], 25)
    m.beta = beta0.astype(np.float64)
    m.expectation_maximization(saving=False)
    out["cut_bounds"] = np.array(m.last_bounds)
    out["cut_final_gamma"] = np.array(m.gamma)
    out["cut_final_sigma"] = np.array(m.sigma)
    print("em_k50: V", V, "cut iterations", len(m.last_bounds), "final bound", m.last_bounds[-1])
    save("em_k50.npz", out)


ALL = dict(
    kat_small=kat_small,
    estep_K5=lambda: estep_fixture("estep_K5.npz", 5, 500, 200, 12345, {0, 2}, 3),
    estep_K20=lambda: estep_fixture("estep_K20.npz", 20, 2000, 96, 1, {0, 1}, 2, keep_m_beta=False),
    estep_K50=lambda: estep_fixture("estep_K50.npz", 50, 2000, 64, 2, {0, 1}, 2, keep_m_beta=False),
    estep_content=estep_content,
    em_c1=em_c1,
    em_c2=em_c2,
    em_k50=em_k50,
    em_toy_ctm=em_toy_ctm,
    wiki_corpus=wiki_corpus,
    spectral=spectral,
    mstep_modes=mstep_modes,
    mnreg=mnreg,
    corpus_params=corpus_params,
)

if __name__ == "__main__":
    names = sys.argv[1:] or list(ALL)
    for n in names:
        ALL[n]()
