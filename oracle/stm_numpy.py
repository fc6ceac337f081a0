"""NumPy/SciPy restatement of the reference's E-step and M-step (TEST INFRASTRUCTURE).

This is the oracle closest to the reference: it calls ``scipy.optimize.minimize(method="BFGS")``
exactly as ``/root/reference/src/modules/stm.py:960-962`` does, so its per-document result goes
through the same third-party state machine (scipy is pinned by the image, identical on the GPU
box).  It is used (a) to validate the C restatement ``stm_oracle.c`` where the live reference is
unavailable, and (b) as the CPU baseline in ``bench.py`` (`cpu_baseline.kind = "port"`), because
its speed is that of the reference's NumPy path.

Pinned bit-for-bit against the live reference by tests/golden/make_golden.py (run in the authoring
container) — see tests/test_oracle_golden.py.
"""
import numpy as np
import scipy.linalg
import scipy.special
from scipy import optimize


# ----------------------------------------------------------------------------------------------
# E-step
# ----------------------------------------------------------------------------------------------

def prologue(sigma):
    """stm.py:497-501: Cholesky of Sigma, entropy term, and the (diagonal) 'siginv' quirk."""
    chol = np.linalg.cholesky(sigma)
    sigmaentropy = np.sum(np.log(np.diag(chol)))
    inv_chol = np.linalg.inv(chol)
    siginv = inv_chol.T * inv_chol  # element-wise: only the diagonal survives
    return siginv, sigmaentropy


def _objective(eta, counts, mu, beta_doc, siginv, K):
    """stm.py:920-944"""
    full = np.insert(eta, K - 1, 0)
    n_tokens = int(np.sum(counts))
    top = full.max()
    prior = 0.5 * (full[:-1] - mu).T @ siginv @ (full[:-1] - mu)
    loglik = np.dot(counts, top + np.log(np.exp(full - top) @ beta_doc))
    return np.float64(prior - (loglik - n_tokens * scipy.special.logsumexp(full)))


def _gradient(eta, counts, mu, beta_doc, siginv, K):
    """stm.py:946-958 (beta not weighted by exp(eta): the reference's quirk, kept)"""
    full = np.insert(eta, K - 1, 0)
    data = beta_doc @ (counts / np.sum(beta_doc.T, axis=1))
    soft = (np.sum(counts) / np.sum(np.exp(full))) * np.exp(full)
    return np.array(np.float64(siginv @ (full[:-1] - mu) - (data - soft)[:-1]))


def _stable_softmax(x):
    """stm.py:905-909"""
    ex = np.exp(x - np.max(x))
    return ex / np.sum(ex)


def _make_pd(M):
    """stm.py:964-984 (in place)"""
    dvec = M.diagonal()
    mags = np.sum(abs(M), 1) - abs(dvec)
    np.fill_diagonal(M, np.where(dvec < mags, mags, dvec))
    return M


def _hessian(eta, counts, beta_doc, siginv, K):
    """stm.py:986-1026 -> (H, repair_stage)"""
    full = np.insert(eta, K - 1, 0)
    theta = _stable_softmax(full)
    a = np.transpose(np.multiply(np.transpose(beta_doc), np.exp(full)))
    b = np.multiply(a, np.transpose(np.sqrt(counts))) / np.sum(a, 0)
    c = np.multiply(b, np.transpose(np.sqrt(counts)))
    hess = b @ b.T - np.sum(counts) * np.multiply(theta[:, None], theta[None, :])
    np.fill_diagonal(hess, np.diag(hess) - np.sum(c, axis=1) + np.sum(counts) * theta)
    H = hess[:-1, :-1] + siginv
    stage = 0
    if not np.all(np.linalg.eigvals(H) > 0):
        H = _make_pd(H)
        stage = 1
        if not np.all(np.linalg.eigvals(H) > 0):
            np.fill_diagonal(H, np.diag(H) + 1e-5)
            stage = 2
    return H, stage


def _decompose(H):
    """stm.py:1031-1050 -> (L, extra_stage)"""
    try:
        return np.linalg.cholesky(H), 0
    except Exception:
        try:
            return np.linalg.cholesky(_make_pd(H)), 4
        except Exception:
            return scipy.linalg.cholesky(_make_pd(H) + 1e-5 * np.eye(H.shape[0])), 12


def infer_document(eta0, mu, counts, beta_doc, siginv, sigmaentropy, K):
    """One pass of the loop body stm.py:519-590 for a single document."""
    res = optimize.minimize(_objective, x0=eta0, args=(counts, mu, beta_doc, siginv, K),
                            jac=_gradient, method="BFGS")
    eta = res.x
    full = np.insert(eta, K - 1, 0)
    theta = np.exp(full) / np.sum(np.exp(full))  # stm.py:547-549, no max shift
    H, stage = _hessian(eta, counts, beta_doc, siginv, K)
    L, stage2 = _decompose(H)
    # lower_bound, stm.py:1068-1101
    sm = _stable_softmax(full)
    diff = eta - mu
    weighted = beta_doc * np.exp(full)[:, None]
    bound = (np.log(sm[None:, ] @ weighted) @ counts - np.sum(np.log(L.diagonal()))
             - 0.5 * diff.T @ siginv @ diff - sigmaentropy)
    # optimize_nu, stm.py:1052-1066
    ui = np.linalg.inv(np.triu(L.T))
    nu = np.dot(ui, np.transpose(ui))
    # update_z, stm.py:1103-1118
    a = np.multiply(beta_doc.T, np.exp(full)).T
    b = np.multiply(a, (np.sqrt(counts) / np.sum(a, 0)))
    phi = np.multiply(b, np.sqrt(counts).T)
    return dict(eta=eta, theta=theta, bound=bound, nu=nu, phi=phi, status=res.status, nit=res.nit,
                nfev=res.nfev, njev=res.njev, repair=stage + stage2, fun=res.fun)


def estep(doc_ptr, word_id, count, beta, mu, siginv, sigmaentropy, eta, aspect=None, docs=None):
    """Full E-step over the CSR corpus (stm.py:489-597); `docs` optionally restricts to a subset
    (accumulators then cover only that subset).  Returns a dict like oracle.c_oracle.estep."""
    beta = np.asarray(beta, dtype=np.float64)
    K = beta.shape[-2]
    D = len(doc_ptr) - 1
    K1 = K - 1
    mu = np.broadcast_to(np.asarray(mu, dtype=np.float64), (D, K1))
    eta = np.array(eta, dtype=np.float64, copy=True).reshape(D, K1)
    out = dict(eta=eta, theta=np.zeros((D, K)), beta_ss=np.zeros(beta.shape),
               sigma_ss=np.zeros((K1, K1)), doc_bound=np.zeros(D), status=np.zeros(D, np.int32),
               nit=np.zeros(D, np.int32), nfev=np.zeros(D, np.int32), njev=np.zeros(D, np.int32),
               repair=np.zeros(D, np.int32))
    bounds = []
    for d in (range(D) if docs is None else docs):
        lo, hi = int(doc_ptr[d]), int(doc_ptr[d + 1])
        ids = np.asarray(word_id[lo:hi], dtype=np.intp)
        cnt = np.asarray(count[lo:hi], dtype=np.float64)
        bmat = beta if aspect is None else beta[int(aspect[d])]
        r = infer_document(eta[d], mu[d], cnt, bmat[:, ids], siginv, sigmaentropy, K)
        eta[d] = r["eta"]
        out["theta"][d] = r["theta"]
        out["sigma_ss"] += r["nu"]
        if aspect is None:
            out["beta_ss"][:, ids] += r["phi"]
        else:
            out["beta_ss"][int(aspect[d])][:, ids] += r["phi"]
        bounds.append(r["bound"])
        for key in ("status", "nit", "nfev", "njev", "repair"):
            out[key][d] = r[key]
        out["doc_bound"][d] = r["bound"]
    out["bound"] = float(np.sum(bounds))
    return out


# ----------------------------------------------------------------------------------------------
# M-step
# ----------------------------------------------------------------------------------------------

def design_matrix(X):
    """stm.py:656-671: covariates as a 2-D array; one-hot encoded column-wise unless exactly 0/1."""
    cov = np.array(X)[:, None]
    if cov.ndim > 2:
        cov = np.squeeze(cov, axis=1)
    if not np.array_equal(cov, cov.astype(bool)):
        cols = []
        for j in range(cov.shape[1]):  # sklearn OneHotEncoder: sorted categories per column
            cats = np.unique(cov[:, j])
            cols.append((cov[:, j][:, None] == cats[None, :]).astype(np.float64))
        cov = np.concatenate(cols, axis=1)
    return np.asarray(cov, dtype=np.float64)


def update_mu(eta, X, model="STM", mode="ols"):
    """stm.py:636-711 -> (mu, gamma). OLS = centred min-norm least squares with the intercept
    dropped afterwards (sklearn LinearRegression.fit: _base.py:702-756, then stm.py:703-706)."""
    N = eta.shape[0]
    if model == "CTM":
        return np.repeat(np.mean(eta, axis=0)[None, :], N, axis=0), None
    if model != "STM":
        raise ValueError("model must be 'STM' or 'CTM'")
    cov = design_matrix(X)
    if mode == "ols" or mode not in ("lasso", "ridge"):
        xc = cov - np.average(cov, axis=0)
        yc = eta - np.average(eta, axis=0)
        coef, _, _, _ = scipy.linalg.lstsq(xc, yc, cond=1e-6)
        gamma = coef.T
    elif mode == "ridge":
        import sklearn.linear_model
        gamma = sklearn.linear_model.Ridge(alpha=0.1, fit_intercept=True).fit(cov, eta).coef_
    else:
        import sklearn.linear_model
        gamma = sklearn.linear_model.Lasso(alpha=1, fit_intercept=True).fit(cov, eta).coef_
    return cov @ gamma.T, gamma


def update_sigma(eta, mu, sigma_ss, sigprior=0.0):
    """stm.py:713-728"""
    resid = eta - mu
    sigma = (resid.T @ resid + sigma_ss) / eta.shape[0]
    return np.diag(np.diag(sigma)) * sigprior + (1 - sigprior) * sigma


def update_beta(beta_ss):
    """stm.py:739-745 (LDA beta: normalise over axis=1, zero-safe)"""
    rs = np.sum(beta_ss, axis=1)[:, None]
    return np.divide(beta_ss, rs, out=np.zeros_like(beta_ss), where=rs != 0)


def em(doc_ptr, word_id, count, beta0, X, n_iter, model="STM", mode="ols", sigprior=0.0,
       threshold=1e-5, aspect=None, estep_fn=None, round_beta32=False, keep_states=False):
    """EM loop (stm.py:855-903) from the reference's initial state (mu=0, eta=0, Sigma=20 I).
    round_beta32: round beta to fp32-representable values after every M-step (what the CUDA path's fp32 beta storage
    does), so that a trace can be compared step by step; keep_states: also return the state every E-step started
    from (beta, mu, sigma, eta) for state-injected ("teacher-forced") comparisons."""
    beta = np.array(beta0, dtype=np.float64)
    if round_beta32:
        beta = beta.astype(np.float32).astype(np.float64)
    K = beta.shape[-2]
    D = len(doc_ptr) - 1
    eta = np.zeros((D, K - 1))
    mu = np.zeros((D, K - 1))
    sigma = np.eye(K - 1) * 20.0
    gamma = None
    bounds, states = [], []
    run = estep if estep_fn is None else estep_fn
    for it in range(100):
        if keep_states:
            states.append(dict(beta=beta.copy(), mu=mu.copy(), sigma=sigma.copy(), eta=eta.copy()))
        siginv, ent = prologue(sigma)
        r = run(doc_ptr, word_id, count, beta, mu, siginv, ent, eta, aspect=aspect)
        eta, theta = r["eta"], r["theta"]
        bounds.append(r["bound"])
        mu, gamma = update_mu(eta, X, model, mode)
        sigma = update_sigma(eta, mu, r["sigma_ss"], sigprior)
        beta = update_beta(r["beta_ss"])
        if round_beta32:
            beta = beta.astype(np.float32).astype(np.float64)
        if it >= 1 and abs((bounds[-1] - bounds[-2]) / abs(bounds[-2])) < threshold:
            break
        if it == n_iter - 1:
            break
    out = dict(beta=beta, theta=theta, eta=eta, mu=mu, sigma=sigma, gamma=gamma, bounds=bounds)
    if keep_states:
        out["states"] = states
    return out


def eval_heldout(doc_ptr, word_id, count, theta, beta, return_doc_ll=False):
    """Held-out likelihood by document completion, restating eval_heldout of
    /root/reference/src/modules/heldout.py:88-97 on CSR arrays: document i is scored with theta[i];
    per-document value = sum_w c_w log(theta_i @ beta[:, w]) / sum_w c_w; result = np.mean over documents."""
    D = len(doc_ptr) - 1
    doc_ll = np.empty(D, dtype=np.float64)
    for i in range(D):
        lo, hi = int(doc_ptr[i]), int(doc_ptr[i + 1])
        w = np.asarray(word_id[lo:hi], dtype=np.intp)
        c = np.asarray(count[lo:hi], dtype=np.float64)
        word_ll = [c[j] * np.log(theta[i] @ beta[:, w[j]]) for j in range(hi - lo)]
        with np.errstate(invalid="ignore", divide="ignore"):
            doc_ll[i] = np.sum(word_ll) / np.sum(c)
    mean = np.mean(doc_ll)
    return (mean, doc_ll) if return_doc_ll else mean


def lasso_from_moments(G, B, yy, N, alpha=1.0, tol=1e-4, max_iter=1000):
    """sklearn Lasso(alpha).fit on centred data, restated on the moments G = Xc'Xc (p x p), B = Xc'Yc
    (p x T), yy[t] = yc_t'yc_t — the form the device M-step evaluates after the all-reduce
    (stm_b200.cu lasso_gamma_kernel).  Follows sklearn/linear_model/_cd_fast.pyx
    enet_coordinate_descent (1.9.0): cyclic coordinate descent, gap-safe screening, duality-gap stop at
    tol * y'y, with X_j'R = b_j - (Gw)_j, R'R = yy - 2 w'b + w'Gw, R'y = yy - w'b.  -> coef (T x p)."""
    p, T = B.shape
    a = alpha * N
    coef = np.zeros((T, p))
    for t in range(T):
        b, w = B[:, t], np.zeros(p)
        excl = np.zeros(p, bool)
        tl = tol * yy[t]

        def gap_enet():
            gw = G @ w
            xta = b - gw
            dn = np.abs(xta).max()
            r2, ry = yy[t] - 2 * w @ b + w @ gw, yy[t] - w @ b
            scale = a / dn if dn > a else 1.0
            return 0.5 * r2 + a * np.abs(w).sum() - (-0.5 * scale ** 2 * r2 + scale * ry), dn, xta

        def screen(first, gap, dn, xta):
            for j in range(p):
                if first:
                    if G[j, j] == 0:
                        w[j], excl[j] = 0.0, True
                        continue
                elif excl[j]:
                    continue
                dj = (1 - abs(xta[j] / max(a, dn))) / np.sqrt(G[j, j])
                if dj <= np.sqrt(2 * gap) / a:
                    excl[j] = False
                else:
                    w[j], excl[j] = 0.0, True

        gap, dn, xta = gap_enet()
        if not gap <= tl:
            screen(True, gap, dn, xta)
            for it in range(max_iter):
                w_max = d_w_max = 0.0
                for j in range(p):
                    if excl[j] or G[j, j] == 0:
                        continue
                    wj = w[j]
                    tmp = (b[j] - G[j] @ w) + wj * G[j, j]
                    w[j] = np.sign(tmp) * max(abs(tmp) - a, 0.0) / G[j, j]
                    d_w_max = max(d_w_max, abs(w[j] - wj))
                    w_max = max(w_max, abs(w[j]))
                if w_max == 0.0 or d_w_max / w_max <= tol or it == max_iter - 1:
                    gap, dn, xta = gap_enet()
                    if gap <= tl:
                        break
                    screen(False, gap, dn, xta)
        coef[t] = w
    return coef
