"""CPU oracle (TEST INFRASTRUCTURE, never imported by the product) for the content-covariate update
`STM.mnreg` — /root/reference/src/modules/stm.py:749-853 ("distributed Poisson regression").

Status of the reference code: as shipped it raises on every supported SciPy (`csr_matrix.A`, stm.py:825, was
removed in SciPy 1.14; the reference pins 1.17.0), hard-codes two aspects (stm.py:762-764, 853) and regresses
EVERY word on column 1 of the count matrix (`counts[:, [1]]`, stm.py:825).  With a one-line compatibility shim
that gives `csr_matrix` its old `.A` property back (tools/ref_shims.py) the unmodified method runs, which is what
tests/golden/mnreg.npz records.  This restatement follows the method line by line and takes the defect as a
parameter:
  * `column=1`     the reference as written — PINNED against the live reference (tests/test_mnreg.py);
  * `column=None`  word i is regressed on ITS OWN column i (what the loop variable is for), aspects general
                   (stack all A blocks, split into A): the specification the product implements by default.
                   This variant has no live-reference run to pin it; it differs from the pinned one only in
                   which column `y` is read from.

What the method computes (stm.py:756-853): counts = vstack(beta_ss[a]) ((A K) x V); covariates per row
r = a K + k: one-hot topic (columns 0..K-1), one-hot aspect (columns K+1..K+A — column K stays empty, the
reference's own "TODO: remove +1"), one-hot interaction (columns K+A+1..K+A+AK); m = log wcounts - log sum
wcounts; per word: sklearn PoissonRegressor(fit_intercept=False, alpha=250, max_iter=1e4, tol=1e-5) of the
column on the covariates — NO offset (`offset2` is computed and never used); kappa = coefficients (p x V);
beta = softmax over the vocabulary of (covar @ kappa + m), split by aspect.
"""
import numpy as np


def covariates(K, A):
    """stm.py:769-793 as a dense (A K) x (K + A + A K + 1) matrix"""
    n = A * K
    X = np.zeros((n, K + A + n + 1))
    r = np.arange(n)
    X[r, np.tile(np.arange(K), A)] = 1.0
    X[r, np.repeat(np.arange(K + 1, K + A + 1), K)] = 1.0
    X[r, np.arange(K + A + 1, K + A + n + 1)] = 1.0
    return X


def poisson_ridge_newton(X, y, alpha, max_iter=100):
    """argmin_w (1/n) sum(exp(Xw) - y * Xw) + alpha/2 |w|^2 — the objective sklearn's PoissonRegressor
    minimises (half Poisson deviance / n + L2), by damped Newton until max|gradient| <= 1e-12 (1 + max y / n).
    Used to check that sklearn's lbfgs answer (gradient tolerance 1e-5) and the device's Newton agree on the
    unique minimiser.  The Armijo test tolerates a decrease below the rounding of f, so that full Newton
    steps continue until the gradient test stops the iteration (same rule as the kernel)."""
    n, p = X.shape
    w = np.zeros(p)
    gtol = 1e-12 * (1.0 + np.abs(y).max() / n)

    def f(w):
        eta = X @ w
        with np.errstate(over="ignore"):
            return np.sum(np.exp(eta) - y * eta) / n + 0.5 * alpha * w @ w

    for _ in range(max_iter):
        mu = np.exp(X @ w)
        g = X.T @ (mu - y) / n + alpha * w
        if np.abs(g).max() <= gtol:
            break
        H = (X.T * mu) @ X / n + alpha * np.eye(p)
        d = -np.linalg.solve(H, g)
        t, f0, slope = 1.0, f(w), g @ d
        while not f(w + t * d) <= f0 + 1e-4 * t * slope + 8.9e-16 * abs(f0) and t > 1e-15:
            t *= 0.5
        w = w + t * d
    return w


def mnreg(beta_ss, wcounts, column=None, solver="sklearn"):
    """-> (beta [A][K][V], kappa (p x V)).  beta_ss: [A][K][V]."""
    beta_ss = np.asarray(beta_ss, dtype=np.float64)
    A, K, V = beta_ss.shape
    counts = beta_ss.reshape(A * K, V)                    # np.concatenate((beta_ss[0], beta_ss[1]), axis=0)
    covar = covariates(K, A)
    m = np.log(wcounts) - np.log(np.sum(wcounts))         # stm.py:795-797 (fixed_intercept)
    out = []
    cache = {}
    for i in range(V):
        col = i if column is None else column
        if col not in cache:
            y = counts[:, col]
            if solver == "sklearn":
                import sklearn.linear_model
                clf = sklearn.linear_model.PoissonRegressor(fit_intercept=False, max_iter=np.intp(1e4), tol=1e-5,
                                                            alpha=np.intp(250))
                cache[col] = clf.fit(covar, y).coef_
            else:
                cache[col] = poisson_ridge_newton(covar, y, 250.0)
        out.append(cache[col])
        if column is None:
            cache.clear()
    coef = np.stack(out, axis=1)                          # p x V
    linpred = m + covar @ coef
    e = np.exp(linpred)
    beta = e / np.sum(e, axis=1)[:, np.newaxis]
    return beta.reshape(A, K, V), coef
