/*
 * stm_oracle.c — CPU fp64 restatement of the reference's per-document Laplace-variational E-step.
 *
 * TEST INFRASTRUCTURE ONLY (see stm_oracle.h).  Every function cites the reference lines it follows:
 *   "stm.py:N"     = /root/reference/src/modules/stm.py
 *   "scipy:F:N"    = site-packages/scipy/optimize/F (scipy 1.18.1 as installed in this image; the
 *                    reference pins 1.17.0 in uv.lock:748-749 — third-party code absent from
 *                    /root/reference, restated here from its published algorithm)
 *
 * The reference's result is defined by the control flow of SciPy's BFGS + MINPACK dcsrch + the
 * Wolfe-2/zoom fallback (its gradient is not the gradient of its objective, SURVEY.md finding 1),
 * so that state machine is restated here branch for branch, including Python's min/max/NaN
 * semantics and the ScalarFunction memoisation that determines nfev/njev.
 */
#include "stm_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <pthread.h>

/* ------------------------------------------------------------------------------------------ */
/* small helpers with Python/NumPy comparison semantics                                        */
/* ------------------------------------------------------------------------------------------ */

/* Python builtin max(a, b, c): keeps the first argument unless a later one compares greater. */
static double py_max3(double a, double b, double c) {
    double m = a;
    if (b > m) m = b;
    if (c > m) m = c;
    return m;
}
static double py_max2(double a, double b) { return (b > a) ? b : a; }
static double py_min2(double a, double b) { return (b < a) ? b : a; }
/* np.clip(x, lo, hi) = minimum(maximum(x, lo), hi), NaN-propagating */
static double np_clip(double x, double lo, double hi) {
    if (isnan(x)) return x;
    double t = (x < lo) ? lo : x;
    return (t > hi) ? hi : t;
}
static double np_sign(double x) {
    if (isnan(x)) return x;
    return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
}

/* ------------------------------------------------------------------------------------------ */
/* per-document problem                                                                        */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    int K, K1, n;
    const double *B;  /* K x n gathered beta columns (stm.py:599-620)            */
    const double *c;  /* n word counts                                            */
    const double *mu; /* K1                                                       */
    const double *S;  /* K1 x K1 siginv                                           */
    double Nsum;      /* np.sum(word_count)         (float, stm.py:955)           */
    double Nint;      /* int(np.sum(word_count))    (truncated, stm.py:933)       */
    double *a;        /* K: eta-independent data term of df (stm.py:954)          */
    double *et, *ex, *d, *w; /* scratch: K, K, K1, max(n, K)                      */
    /* ScalarFunction memo (scipy:_differentiable_functions.py:337-401)            */
    double *xc, *gc;
    double fc;
    int have_x, f_ok, g_ok;
    int nfev, njev;
} doc_t;

/* objective, stm.py:920-944 */
static double obj_f(doc_t *p, const double *eta) {
    const int K = p->K, K1 = p->K1, n = p->n;
    double *et = p->et, *ex = p->ex, *d = p->d;
    double m = 0.0; /* eta~ = [eta, 0]  (np.insert(eta, K-1, 0)) */
    for (int k = 0; k < K1; ++k) { et[k] = eta[k]; d[k] = eta[k] - p->mu[k]; }
    et[K1] = 0.0;
    m = et[0];
    for (int k = 1; k < K; ++k) if (et[k] > m) m = et[k];
    if (isnan(m)) m = NAN;
    for (int k = 0; k < K; ++k) if (isnan(et[k])) m = NAN; /* np.max propagates NaN */

    /* 0.5 * d' S d : (d @ S) @ d */
    double quad = 0.0;
    for (int j = 0; j < K1; ++j) {
        double t = 0.0;
        for (int i = 0; i < K1; ++i) t += d[i] * p->S[(size_t)i * K1 + j];
        quad += t * d[j];
    }
    quad *= 0.5;

    /* np.dot(c, m + log(exp(eta~ - m) @ B)) */
    for (int k = 0; k < K; ++k) ex[k] = exp(et[k] - m);
    double data = 0.0;
    for (int v = 0; v < n; ++v) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += ex[k] * p->B[(size_t)k * n + v];
        data += p->c[v] * (m + log(s));
    }

    /* scipy.special.logsumexp(eta~)  (scipy/special/_logsumexp.py:201-247):
     * elements equal to the max are pulled out of the sum: log1p(s/cnt) + log(cnt) + max */
    double cnt = 0.0, s = 0.0;
    for (int k = 0; k < K; ++k) {
        if (et[k] == m) cnt += 1.0; else s += exp(et[k] - m);
    }
    if (s != 0.0) s = s / cnt;
    double lse = log1p(s) + log(cnt) + m;
    if (!isfinite(lse)) { /* wrapper falls back to log(sum(exp(a))) (_logsumexp.py:117-134) */
        double t = 0.0;
        for (int k = 0; k < K; ++k) t += exp(et[k]);
        lse = log(t);
    }
    return quad - (data - p->Nint * lse);
}

/* "gradient", stm.py:946-958 — beta is NOT weighted by exp(eta) (reference quirk) */
static void obj_df(doc_t *p, const double *eta, double *g) {
    const int K = p->K, K1 = p->K1;
    double *et = p->et, *ex = p->ex, *d = p->d;
    for (int k = 0; k < K1; ++k) { et[k] = eta[k]; d[k] = eta[k] - p->mu[k]; }
    et[K1] = 0.0;
    double se = 0.0;
    for (int k = 0; k < K; ++k) { ex[k] = exp(et[k]); se += ex[k]; } /* no max shift, stm.py:955 */
    double scale = p->Nsum / se;
    for (int i = 0; i < K1; ++i) {
        double t = 0.0;
        for (int j = 0; j < K1; ++j) t += p->S[(size_t)i * K1 + j] * d[j];
        g[i] = t - (p->a[i] - scale * ex[i]);
    }
}

/* a_k = sum_v B_kv * (c_v / sum_k' B_k'v)   — stm.py:954, eta-independent */
static void precompute_a(doc_t *p) {
    const int K = p->K, n = p->n;
    double *r = p->w;
    for (int v = 0; v < n; ++v) {
        double cs = 0.0;
        for (int k = 0; k < K; ++k) cs += p->B[(size_t)k * n + v];
        r[v] = p->c[v] / cs;
    }
    for (int k = 0; k < K; ++k) {
        double t = 0.0;
        for (int v = 0; v < n; ++v) t += p->B[(size_t)k * n + v] * r[v];
        p->a[k] = t;
    }
}

/* Optional self-check of the shortcut the CUDA kernel takes (estep_kernel.cuh, STM_CURV_CERT).  The reference's
 * gradient is the gradient of a CONVEX function h (the data term of df is not weighted by exp(eta), stm.py:954), so
 * phi'(alpha) = g(x + alpha p).p cannot rise by more than alpha*C, C = p'Sp + N min(max p_k^2, |p|^2/2) (or
 * p'Sp + 1.11 N Var_theta(x)([p,0]) while alpha (max pt - min pt) <= 0.1), and for alpha <= a_safe = 0.09|phi'(0)|/C
 * the strong-Wolfe curvature test |phi'(alpha)| <= 0.9|phi'(0)| - part of every acceptance test - cannot pass.  The
 * kernel ends DCSRCH (bracket set) and _zoom (a_lo <= a_hi) as soon as their bracket lies inside [0, a_safe].  With the
 * check enabled the oracle keeps replaying every search in full, as SciPy does, evaluates the same rule on the way, and
 * counts the searches in which the certificate held and those of them that went on to ACCEPT a step (must stay 0),
 * plus the trials the kernel does not make. */
static int shortcut_check_on = 0;
static long long cert_w1_fired = 0, cert_w1_accept_after = 0, cert_zoom_fired = 0, cert_zoom_accept_after = 0;
static long long cert_trials_skipped = 0;
void stm_oracle_shortcut_check(int enable, long long *out5) {
    if (out5) {
        out5[0] = cert_w1_fired; out5[1] = cert_w1_accept_after; out5[2] = cert_zoom_fired;
        out5[3] = cert_zoom_accept_after; out5[4] = cert_trials_skipped;
    }
    shortcut_check_on = enable;
    cert_w1_fired = cert_w1_accept_after = cert_zoom_fired = cert_zoom_accept_after = cert_trials_skipped = 0;
}

static int vec_equal(const double *a, const double *b, int n) {
    for (int i = 0; i < n; ++i) if (!(a[i] == b[i])) return 0;
    return 1;
}
/* ScalarFunction.fun / .grad with memoisation on x (decides nfev / njev) */
static void memo_set_x(doc_t *p, const double *x) {
    if (!p->have_x || !vec_equal(x, p->xc, p->K1)) {
        memcpy(p->xc, x, sizeof(double) * p->K1);
        p->have_x = 1; p->f_ok = 0; p->g_ok = 0;
    }
}
static double sf_fun(doc_t *p, const double *x) {
    memo_set_x(p, x);
    if (!p->f_ok) { p->fc = obj_f(p, p->xc); p->f_ok = 1; p->nfev++; }
    return p->fc;
}
static const double *sf_grad(doc_t *p, const double *x) {
    memo_set_x(p, x);
    if (!p->g_ok) { obj_df(p, p->xc, p->gc); p->g_ok = 1; p->njev++; }
    return p->gc;
}

/* ------------------------------------------------------------------------------------------ */
/* line searches                                                                               */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    doc_t *p;
    const double *xk, *pk;
    double *xt;   /* K1 trial point                    */
    double *gval; /* K1 gradient at last derphi() call */
    int have_gval;
    double a_safe; /* curvature certificate of the current direction (0 = none); only used by the self-check */
} line_t;

static double ls_phi(line_t *L, double s) {
    for (int i = 0; i < L->p->K1; ++i) L->xt[i] = L->xk[i] + s * L->pk[i];
    return sf_fun(L->p, L->xt);
}
static double ls_derphi(line_t *L, double s) {
    const int n = L->p->K1;
    for (int i = 0; i < n; ++i) L->xt[i] = L->xk[i] + s * L->pk[i];
    const double *g = sf_grad(L->p, L->xt);
    memcpy(L->gval, g, sizeof(double) * n);
    L->have_gval = 1;
    double t = 0.0;
    for (int i = 0; i < n; ++i) t += g[i] * L->pk[i];
    return t;
}

/* MINPACK-2 dcstep — scipy:_dcsrch.py:502-728 */
static void dcstep(double *stx, double *fx, double *dx, double *sty, double *fy, double *dy,
                   double *stp, double fp, double dp, int *brackt, double stpmin, double stpmax) {
    double sgnd = np_sign(dp) * np_sign(*dx);
    double theta, s, gamma, p, q, r, stpc, stpq, stpf;

    if (fp > *fx) {
        theta = 3.0 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = py_max3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp < *stx) gamma *= -1;
        p = (gamma - *dx) + theta;
        q = ((gamma - *dx) + gamma) + dp;
        r = p / q;
        stpc = *stx + r * (*stp - *stx);
        stpq = *stx + ((*dx / ((*fx - fp) / (*stp - *stx) + *dx)) / 2.0) * (*stp - *stx);
        if (fabs(stpc - *stx) <= fabs(stpq - *stx)) stpf = stpc;
        else stpf = stpc + (stpq - stpc) / 2.0;
        *brackt = 1;
    } else if (sgnd < 0.0) {
        theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = py_max3(fabs(theta), fabs(*dx), fabs(dp));
        gamma = s * sqrt((theta / s) * (theta / s) - (*dx / s) * (dp / s));
        if (*stp > *stx) gamma *= -1;
        p = (gamma - dp) + theta;
        q = ((gamma - dp) + gamma) + *dx;
        r = p / q;
        stpc = *stp + r * (*stx - *stp);
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
        else stpf = stpq;
        *brackt = 1;
    } else if (fabs(dp) < fabs(*dx)) {
        theta = 3 * (*fx - fp) / (*stp - *stx) + *dx + dp;
        s = py_max3(fabs(theta), fabs(*dx), fabs(dp));
        /* max(0, x): Python builtin, keeps 0 when x is NaN */
        gamma = s * sqrt(py_max2(0.0, (theta / s) * (theta / s) - (*dx / s) * (dp / s)));
        if (*stp > *stx) gamma = -gamma;
        p = (gamma - dp) + theta;
        q = (gamma + (*dx - dp)) + gamma;
        r = p / q;
        if (r < 0 && gamma != 0) stpc = *stp + r * (*stx - *stp);
        else if (*stp > *stx) stpc = stpmax;
        else stpc = stpmin;
        stpq = *stp + (dp / (dp - *dx)) * (*stx - *stp);
        if (*brackt) {
            if (fabs(stpc - *stp) < fabs(stpq - *stp)) stpf = stpc;
            else stpf = stpq;
            if (*stp > *stx) stpf = py_min2(*stp + 0.66 * (*sty - *stp), stpf);
            else stpf = py_max2(*stp + 0.66 * (*sty - *stp), stpf);
        } else {
            if (fabs(stpc - *stp) > fabs(stpq - *stp)) stpf = stpc;
            else stpf = stpq;
            stpf = np_clip(stpf, stpmin, stpmax);
        }
    } else {
        if (*brackt) {
            theta = 3.0 * (fp - *fy) / (*sty - *stp) + *dy + dp;
            s = py_max3(fabs(theta), fabs(*dy), fabs(dp));
            gamma = s * sqrt((theta / s) * (theta / s) - (*dy / s) * (dp / s));
            if (*stp > *sty) gamma = -gamma;
            p = (gamma - dp) + theta;
            q = ((gamma - dp) + gamma) + *dy;
            r = p / q;
            stpc = *stp + r * (*sty - *stp);
            stpf = stpc;
        } else if (*stp > *stx) stpf = stpmax;
        else stpf = stpmin;
    }

    if (fp > *fx) {
        *sty = *stp; *fy = fp; *dy = dp;
    } else {
        if (sgnd < 0) { *sty = *stx; *fy = *fx; *dy = *dx; }
        *stx = *stp; *fx = fp; *dx = dp;
    }
    *stp = stpf;
}

/* scalar_search_wolfe1 + DCSRCH.__call__ and _iterate — scipy:_linesearch.py:108-178, _dcsrch.py:201-500.
 * Returns 1 with alpha and phi1 set on success, 0 on failure. */
static int search_wolfe1(line_t *L, double phi0, double old_phi0, double derphi0,
                         double *alpha, double *phi1_out) {
    const double ftol = 1e-4, gtol = 0.9, xtol = 1e-14, stpmin = 1e-100, stpmax = 1e100;
    double alpha1;
    if (derphi0 != 0) {
        alpha1 = py_min2(1.0, 1.01 * 2 * (phi0 - old_phi0) / derphi0);
        if (alpha1 < 0) alpha1 = 1.0;
    } else alpha1 = 1.0;

    /* START (_dcsrch.py:264-308) */
    double stp = alpha1, f = phi0, g = derphi0;
    if (stp < stpmin || stp > stpmax || g >= 0) return 0; /* task = ERROR */
    int brackt = 0, stage = 1;
    double finit = f, ginit = g, gtest = ftol * ginit;
    double width = stpmax - stpmin, width1 = width / 0.5;
    double stx = 0.0, fx = finit, gx = ginit, sty = 0.0, fy = finit, gy = ginit;
    double stmin = 0, stmax = stp + 4.0 * stp;
    int certw = 0;

    /* DCSRCH.__call__: for i in range(maxiter=100); i = 0 was START */
    for (int it = 0; it < 100; ++it) {
        if (it > 0) {
            /* one _iterate call with (stp, f, g) */
            double ftest = finit + stp * gtest;
            int warn = 0, conv = 0;
            if (stage == 1 && f <= ftest && g >= 0) stage = 2;
            if (brackt && (stp <= stmin || stp >= stmax)) warn = 1;
            if (brackt && stmax - stmin <= xtol * stmax) warn = 1;
            if (stp == stpmax && f <= ftest && g <= gtest) warn = 1;
            if (stp == stpmin && (f > ftest || g >= gtest)) warn = 1;
            if (f <= ftest && fabs(g) <= gtol * -ginit) conv = 1;
            if (conv) {
                if (certw) __sync_fetch_and_add(&cert_w1_accept_after, 1);
                *alpha = stp; *phi1_out = f; return isfinite(stp) ? 1 : 0;
            }
            if (warn) return 0;

            if (stage == 1 && f <= fx && f > ftest) {
                double fm = f - stp * gtest, fxm = fx - stx * gtest, fym = fy - sty * gtest;
                double gm = g - gtest, gxm = gx - gtest, gym = gy - gtest;
                dcstep(&stx, &fxm, &gxm, &sty, &fym, &gym, &stp, fm, gm, &brackt, stmin, stmax);
                fx = fxm + stx * gtest; fy = fym + sty * gtest;
                gx = gxm + gtest; gy = gym + gtest;
            } else {
                dcstep(&stx, &fx, &gx, &sty, &fy, &gy, &stp, f, g, &brackt, stmin, stmax);
            }
            if (brackt) {
                if (fabs(sty - stx) >= 0.66 * width1) stp = stx + 0.5 * (sty - stx);
                width1 = width;
                width = fabs(sty - stx);
            }
            if (brackt) {
                stmin = py_min2(stx, sty);
                stmax = py_max2(stx, sty);
            } else {
                stmin = stp + 1.1 * (stp - stx);
                stmax = stp + 4.0 * (stp - stx);
            }
            stp = np_clip(stp, stpmin, stpmax);
            if ((brackt && (stp <= stmin || stp >= stmax)) ||
                (brackt && stmax - stmin <= xtol * stmax))
                stp = stx;
            if (shortcut_check_on && brackt && !certw && L->a_safe > 0.0 && stmax <= L->a_safe) {
                certw = 1;
                __sync_fetch_and_add(&cert_w1_fired, 1);
            }
        }
        if (!isfinite(stp)) return 0;
        if (certw) __sync_fetch_and_add(&cert_trials_skipped, 1);
        /* task == FG */
        f = ls_phi(L, stp);
        g = ls_derphi(L, stp);
    }
    return 0; /* maxiter reached */
}

/* scipy:_linesearch.py:491-522; NaN stands for None */
static double cubicmin(double a, double fa, double fpa, double b, double fb, double c, double fc) {
    double C = fpa, db = b - a, dc = c - a;
    double denom = (db * dc) * (db * dc) * (db - dc);
    double d00 = dc * dc, d01 = -(db * db), d10 = -(dc * dc * dc), d11 = db * db * db;
    double v0 = fb - fa - C * db, v1 = fc - fa - C * dc;
    double A = d00 * v0 + d01 * v1;
    double B = d10 * v0 + d11 * v1;
    if (denom == 0.0) return NAN; /* divide='raise' */
    A /= denom;
    B /= denom;
    double radical = B * B - 3 * A * C;
    if (radical < 0.0 || A == 0.0) return NAN; /* invalid / divide raise */
    double xmin = a + (-B + sqrt(radical)) / (3 * A);
    if (!isfinite(xmin)) return NAN;
    return xmin;
}
/* scipy:_linesearch.py:525-543 */
static double quadmin(double a, double fa, double fpa, double b, double fb) {
    double D = fa, C = fpa, db = b - a * 1.0;
    if (db * db == 0.0) return NAN;
    double B = (fb - D - C * db) / (db * db);
    if (2.0 * B == 0.0) return NAN;
    double xmin = a - C / (2.0 * B);
    if (!isfinite(xmin)) return NAN;
    return xmin;
}

/* _zoom — scipy:_linesearch.py:546-634. returns 1 on success */
static int zoom(line_t *L, double a_lo, double a_hi, double phi_lo, double phi_hi, double derphi_lo,
                double phi0, double derphi0, double c1, double c2,
                double *a_star, double *val_star) {
    const int maxiter = 10;
    int i = 0;
    const double delta1 = 0.2, delta2 = 0.1;
    double phi_rec = phi0, a_rec = 0;
    double a_j = NAN;
    int certz = 0;
    for (;;) {
        if (shortcut_check_on && !certz && L->a_safe > 0.0 && a_lo >= 0.0 && a_lo <= a_hi && a_hi <= L->a_safe) {
            certz = 1;
            __sync_fetch_and_add(&cert_zoom_fired, 1);
        }
        if (certz) __sync_fetch_and_add(&cert_trials_skipped, 1);
        double dalpha = a_hi - a_lo, a, b, cchk = 0.0;
        if (dalpha < 0) { a = a_hi; b = a_lo; } else { a = a_lo; b = a_hi; }
        if (i > 0) {
            cchk = delta1 * dalpha;
            a_j = cubicmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi, a_rec, phi_rec);
        }
        if (i == 0 || isnan(a_j) || a_j > b - cchk || a_j < a + cchk) {
            double qchk = delta2 * dalpha;
            a_j = quadmin(a_lo, phi_lo, derphi_lo, a_hi, phi_hi);
            if (isnan(a_j) || a_j > b - qchk || a_j < a + qchk) a_j = a_lo + 0.5 * dalpha;
        }
        double phi_aj = ls_phi(L, a_j);
        if (phi_aj > phi0 + c1 * a_j * derphi0 || phi_aj >= phi_lo) {
            phi_rec = phi_hi; a_rec = a_hi; a_hi = a_j; phi_hi = phi_aj;
        } else {
            double derphi_aj = ls_derphi(L, a_j);
            if (fabs(derphi_aj) <= -c2 * derphi0) {
                if (certz) __sync_fetch_and_add(&cert_zoom_accept_after, 1);
                *a_star = a_j; *val_star = phi_aj; return 1;
            }
            if (derphi_aj * (a_hi - a_lo) >= 0) {
                phi_rec = phi_hi; a_rec = a_hi; a_hi = a_lo; phi_hi = phi_lo;
            } else {
                phi_rec = phi_lo; a_rec = a_lo;
            }
            a_lo = a_j; phi_lo = phi_aj; derphi_lo = derphi_aj;
        }
        i += 1;
        if (i > maxiter) return 0;
    }
}

/* scalar_search_wolfe2 — scipy:_linesearch.py:343-488 (amax = 1e100, maxiter = 10).
 * returns 0 fail; 1 success with gradient at alpha in L->gval; 2 success without gradient
 * (bracket phase ran out: derphi_star None). */
static int search_wolfe2(line_t *L, double phi0, double old_phi0, double derphi0,
                         double *alpha, double *phi_star) {
    const double c1 = 1e-4, c2 = 0.9, amax = 1e100;
    double alpha0 = 0, alpha1;
    if (derphi0 != 0) alpha1 = py_min2(1.0, 1.01 * 2 * (phi0 - old_phi0) / derphi0);
    else alpha1 = 1.0;
    if (alpha1 < 0) alpha1 = 1.0;
    alpha1 = py_min2(alpha1, amax);

    double phi_a1 = ls_phi(L, alpha1);
    double phi_a0 = phi0, derphi_a0 = derphi0;

    for (int i = 0; i < 10; ++i) {
        if (alpha1 == 0 || alpha0 > amax) return 0;
        if (phi_a1 > phi0 + c1 * alpha1 * derphi0 || (phi_a1 >= phi_a0 && i > 0))
            return zoom(L, alpha0, alpha1, phi_a0, phi_a1, derphi_a0, phi0, derphi0, c1, c2,
                        alpha, phi_star);
        double derphi_a1 = ls_derphi(L, alpha1);
        if (fabs(derphi_a1) <= -c2 * derphi0) { *alpha = alpha1; *phi_star = phi_a1; return 1; }
        if (derphi_a1 >= 0)
            return zoom(L, alpha1, alpha0, phi_a1, phi_a0, derphi_a1, phi0, derphi0, c1, c2,
                        alpha, phi_star);
        double alpha2 = 2 * alpha1;
        alpha2 = py_min2(alpha2, amax);
        alpha0 = alpha1; alpha1 = alpha2;
        phi_a0 = phi_a1;
        phi_a1 = ls_phi(L, alpha1);
        derphi_a0 = derphi_a1;
    }
    *alpha = alpha1; *phi_star = phi_a1;
    return 2;
}

/* ------------------------------------------------------------------------------------------ */
/* _minimize_bfgs — scipy:_optimize.py:1345-1526                                               */
/* ------------------------------------------------------------------------------------------ */

static int bfgs(doc_t *p, double *x, double *work /* 6*n + 3*n*n */, double *fun_out, int *nit_out) {
    const int n = p->K1;
    double *gfk = work, *pk = gfk + n, *sk = pk + n, *yk = sk + n, *xt = yk + n, *gval = xt + n;
    double *Hk = gval + n, *T1 = Hk + (size_t)n * n, *T2 = T1 + (size_t)n * n;
    const int maxiter = n * 200;
    const double gtol = 1e-5;

    double old_fval = sf_fun(p, x);
    memcpy(gfk, sf_grad(p, x), sizeof(double) * n);
    int k = 0, warnflag = 0;
    for (int i = 0; i < n; ++i)
        for (int j = 0; j < n; ++j) Hk[(size_t)i * n + j] = (i == j) ? 1.0 : 0.0;
    double nrm2 = 0.0;
    for (int i = 0; i < n; ++i) nrm2 += gfk[i] * gfk[i];
    double old_old_fval = old_fval + sqrt(nrm2) / 2;
    double gnorm = 0.0;
    for (int i = 0; i < n; ++i) { double t = fabs(gfk[i]); if (t > gnorm || isnan(t)) gnorm = t; }

    line_t L;
    L.p = p; L.xk = x; L.pk = pk; L.xt = xt; L.gval = gval; L.have_gval = 0;

    while (gnorm > gtol && k < maxiter) {
        for (int i = 0; i < n; ++i) {
            double t = 0.0;
            for (int j = 0; j < n; ++j) t += Hk[(size_t)i * n + j] * gfk[j];
            pk[i] = -t;
        }
        double derphi0 = 0.0;
        for (int i = 0; i < n; ++i) derphi0 += gfk[i] * pk[i];

        L.a_safe = 0.0;
        if (shortcut_check_on) {
            /* the kernel's formula and guard (estep_kernel.cuh, "curvature certificate"): bound (i) on the variance
             * term, and bound (ii) from theta at x */
            double pSp = 0.0, mx = 0.0, sm = 0.0, noise = 0.0, hi = 0.0, lo = 0.0, xm = 0.0;
            for (int i = 0; i < n; ++i) {
                double t = 0.0, sx = 0.0;
                for (int j = 0; j < n; ++j) { t += p->S[(size_t)i * n + j] * pk[j]; sx += p->S[(size_t)i * n + j] * (x[j] - p->mu[j]); }
                pSp += pk[i] * t;
                if (pk[i] * pk[i] > mx) mx = pk[i] * pk[i];
                if (pk[i] > hi) hi = pk[i];
                if (-pk[i] > lo) lo = -pk[i];
                if (x[i] > xm) xm = x[i];
                sm += pk[i] * pk[i];
                noise += fabs(pk[i]) * (fabs(sx) + fabs(p->a[i]) + p->Nsum);
            }
            const double C1 = pSp + p->Nsum * fmin(mx, 0.5 * sm);
            if (derphi0 < 0.0 && C1 > 0.0 && C1 < 1e300 && 1e-12 * noise <= 0.01 * -derphi0) {
                const double num = 0.09 * -derphi0;
                double as = num / C1;
                double se = exp(0.0 - xm), m1 = 0.0, v0 = 0.0;
                for (int i = 0; i < n; ++i) se += exp(x[i] - xm);
                for (int i = 0; i < n; ++i) m1 += exp(x[i] - xm) / se * pk[i];
                for (int i = 0; i < n; ++i) v0 += exp(x[i] - xm) / se * (pk[i] - m1) * (pk[i] - m1);
                v0 += exp(0.0 - xm) / se * m1 * m1;
                const double C2 = pSp + 1.1100001 * p->Nsum * v0, R = hi + lo;
                if (C2 > 0.0 && v0 >= 0.0 && R > 0.0 && R < 1e300) {
                    const double as2 = fmin(num / C2, 0.1 / R);
                    if (as2 > as) as = as2;
                }
                if (isfinite(as)) L.a_safe = as;
            }
        }
        double alpha_k = 0.0, new_fval = 0.0;
        int have_g = 0;
        int ok = search_wolfe1(&L, old_fval, old_old_fval, derphi0, &alpha_k, &new_fval);
        if (ok) have_g = 1;
        else {
            ok = search_wolfe2(&L, old_fval, old_old_fval, derphi0, &alpha_k, &new_fval);
            have_g = (ok == 1);
        }
        if (!ok) { warnflag = 2; break; }
        old_old_fval = old_fval;
        old_fval = new_fval;

        for (int i = 0; i < n; ++i) { sk[i] = alpha_k * pk[i]; x[i] = x[i] + sk[i]; }
        if (!have_g) memcpy(gval, sf_grad(p, x), sizeof(double) * n);
        for (int i = 0; i < n; ++i) { yk[i] = gval[i] - gfk[i]; gfk[i] = gval[i]; }
        k += 1;
        gnorm = 0.0;
        for (int i = 0; i < n; ++i) { double t = fabs(gfk[i]); if (t > gnorm || isnan(t)) gnorm = t; }
        if (gnorm <= gtol) break;
        /* xrtol = 0: alpha*|pk|_inf <= 0 */
        double pn = 0.0;
        for (int i = 0; i < n; ++i) { double t = fabs(pk[i]); if (t > pn || isnan(t)) pn = t; }
        if (alpha_k * pn <= 0.0) break;
        if (!isfinite(old_fval)) { warnflag = 2; break; }

        double rhok_inv = 0.0, rhok;
        for (int i = 0; i < n; ++i) rhok_inv += yk[i] * sk[i];
        rhok = (rhok_inv == 0.0) ? 1000.0 : 1.0 / rhok_inv;
        /* Hk = A1 @ (Hk @ A2) + rhok*sk*sk',  A1 = I - sk*yk'*rhok, A2 = I - yk*sk'*rhok */
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double t = 0.0;
                for (int l = 0; l < n; ++l) {
                    double a2 = ((l == j) ? 1.0 : 0.0) - yk[l] * sk[j] * rhok;
                    t += Hk[(size_t)i * n + l] * a2;
                }
                T1[(size_t)i * n + j] = t;
            }
        for (int i = 0; i < n; ++i)
            for (int j = 0; j < n; ++j) {
                double t = 0.0;
                for (int l = 0; l < n; ++l) {
                    double a1 = ((i == l) ? 1.0 : 0.0) - sk[i] * yk[l] * rhok;
                    t += a1 * T1[(size_t)l * n + j];
                }
                T2[(size_t)i * n + j] = t + rhok * sk[i] * sk[j];
            }
        memcpy(Hk, T2, sizeof(double) * n * n);
    }

    if (warnflag != 2) {
        if (k >= maxiter) warnflag = 1;
        else {
            int anynan = isnan(gnorm) || isnan(old_fval);
            for (int i = 0; i < n; ++i) anynan |= isnan(x[i]);
            warnflag = anynan ? 3 : 0;
        }
    }
    *fun_out = old_fval;
    *nit_out = k;
    return warnflag;
}

/* ------------------------------------------------------------------------------------------ */
/* dense helpers: Cholesky (lower), PD test, make_pd                                           */
/* ------------------------------------------------------------------------------------------ */

/* lower Cholesky of the symmetric n x n matrix M into Lo (only lower part referenced); 0 ok, -1 not PD */
static int chol_lower(int n, const double *M, double *Lo) {
    memset(Lo, 0, sizeof(double) * n * n);
    for (int j = 0; j < n; ++j) {
        double s = M[(size_t)j * n + j];
        for (int k = 0; k < j; ++k) s -= Lo[(size_t)j * n + k] * Lo[(size_t)j * n + k];
        if (!(s > 0.0)) return -1;
        double ljj = sqrt(s);
        Lo[(size_t)j * n + j] = ljj;
        for (int i = j + 1; i < n; ++i) {
            double t = M[(size_t)i * n + j];
            for (int k = 0; k < j; ++k) t -= Lo[(size_t)i * n + k] * Lo[(size_t)j * n + k];
            Lo[(size_t)i * n + j] = t / ljj;
        }
    }
    return 0;
}

/* make_pd — stm.py:964-984 (in place) */
static void make_pd(int n, double *M) {
    for (int i = 0; i < n; ++i) {
        double tot = 0.0;
        for (int j = 0; j < n; ++j) tot += fabs(M[(size_t)i * n + j]);
        double dv = M[(size_t)i * n + i];
        double mag = tot - fabs(dv);
        if (dv < mag) M[(size_t)i * n + i] = mag;
    }
}

int stm_oracle_prologue(int K1, const double *sigma, double *siginv, double *sigmaentropy) {
    double *Lo = (double *)malloc(sizeof(double) * K1 * K1);
    if (chol_lower(K1, sigma, Lo) != 0) { free(Lo); return -1; }
    double ent = 0.0;
    memset(siginv, 0, sizeof(double) * K1 * K1);
    for (int i = 0; i < K1; ++i) {
        double l = Lo[(size_t)i * K1 + i];
        ent += log(l);
        double il = 1.0 / l; /* inv(L)_ii; off-diagonals vanish in inv(L).T * inv(L), stm.py:501 */
        siginv[(size_t)i * K1 + i] = il * il;
    }
    *sigmaentropy = ent;
    free(Lo);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* per-document pipeline — stm.py:519-590                                                      */
/* ------------------------------------------------------------------------------------------ */

typedef struct {
    double bound;
    int status, nit, nfev, njev, repair;
} doc_out_t;

static size_t doc_work_size(int K, int nmax) {
    const size_t K1 = K - 1;
    size_t n = 0;
    n += (size_t)K * nmax;           /* B   */
    n += (size_t)K * nmax;           /* phi / b scratch */
    n += nmax;                       /* c   */
    n += (size_t)K;                  /* a   */
    n += (size_t)K * 2 + K1;         /* et, ex, d */
    n += (size_t)(nmax > K ? nmax : K); /* w */
    n += K1 * 2;                     /* xc, gc */
    n += 6 * K1 + 3 * K1 * K1;       /* bfgs work */
    n += 3 * K1 * K1 + (size_t)K * K; /* H, L, Linv, hess full */
    n += (size_t)K * 2;              /* theta, tmp */
    return n;
}

/* One document. phi (K x n) and nu (K1 x K1) are returned in caller buffers for in-order accumulation. */
static void infer_doc(int K, int n, const int32_t *wid, const double *cnt, const double *beta_a /*K x V*/,
                      int V, const double *mu, const double *S, double sigmaentropy,
                      double *eta, double *theta_out, double *work, double *phi_out, double *nu_out,
                      doc_out_t *out) {
    const int K1 = K - 1;
    double *w = work;
    double *B = w; w += (size_t)K * n;
    double *bq = w; w += (size_t)K * n;
    double *c = w; w += n;
    doc_t P;
    P.K = K; P.K1 = K1; P.n = n; P.B = B; P.c = c; P.mu = mu; P.S = S;
    P.a = w; w += K;
    P.et = w; w += K;
    P.ex = w; w += K;
    P.d = w; w += K1;
    P.w = w; w += (n > K ? n : K);
    P.xc = w; w += K1;
    P.gc = w; w += K1;
    double *bw = w; w += 6 * (size_t)K1 + 3 * (size_t)K1 * K1;
    double *H = w; w += (size_t)K1 * K1;
    double *Lo = w; w += (size_t)K1 * K1;
    double *Li = w; w += (size_t)K1 * K1;
    double *HF = w; w += (size_t)K * K;
    double *th = w; w += K;
    double *tmp = w; w += K;
    P.have_x = 0; P.f_ok = 0; P.g_ok = 0; P.nfev = 0; P.njev = 0;

    /* gather — get_beta, stm.py:599-620 */
    double Nsum = 0.0;
    for (int v = 0; v < n; ++v) { c[v] = cnt[v]; Nsum += cnt[v]; }
    for (int k = 0; k < K; ++k)
        for (int v = 0; v < n; ++v) B[(size_t)k * n + v] = beta_a[(size_t)k * V + wid[v]];
    P.Nsum = Nsum;
    P.Nint = (double)(long long)Nsum;
    precompute_a(&P);

    /* optimize_eta — stm.py:917-962 */
    double fun;
    out->status = bfgs(&P, eta, bw, &fun, &out->nit);
    out->nfev = P.nfev; out->njev = P.njev;

    /* theta — stm.py:547-549 (no max shift) */
    double *et = P.et, *ex = P.ex;
    for (int k = 0; k < K1; ++k) et[k] = eta[k];
    et[K1] = 0.0;
    double se = 0.0;
    for (int k = 0; k < K; ++k) { ex[k] = exp(et[k]); se += ex[k]; }
    for (int k = 0; k < K; ++k) theta_out[k] = ex[k] / se;

    /* stable softmax used by hessian() and lower_bound() — stm.py:905-909 */
    double m = et[0];
    for (int k = 1; k < K; ++k) if (et[k] > m) m = et[k];
    double ss = 0.0;
    for (int k = 0; k < K; ++k) { tmp[k] = exp(et[k] - m); ss += tmp[k]; }
    for (int k = 0; k < K; ++k) th[k] = tmp[k] / ss;

    /* hessian — stm.py:986-1026.  a = B o exp(eta~); b = a*sqrt(c)/colsum(a); cmat = b*sqrt(c) */
    double *colsum = P.w;
    for (int v = 0; v < n; ++v) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += B[(size_t)k * n + v] * ex[k];
        colsum[v] = s;
    }
    for (int k = 0; k < K; ++k)
        for (int v = 0; v < n; ++v)
            bq[(size_t)k * n + v] = (B[(size_t)k * n + v] * ex[k]) * sqrt(c[v]) / colsum[v];
    for (int i = 0; i < K; ++i)
        for (int j = 0; j <= i; ++j) {
            double t = 0.0;
            for (int v = 0; v < n; ++v) t += bq[(size_t)i * n + v] * bq[(size_t)j * n + v];
            t -= Nsum * (th[i] * th[j]);
            HF[(size_t)i * K + j] = t;
            HF[(size_t)j * K + i] = t;
        }
    for (int k = 0; k < K; ++k) {
        double rs = 0.0;
        for (int v = 0; v < n; ++v) rs += bq[(size_t)k * n + v] * sqrt(c[v]);
        HF[(size_t)k * K + k] = HF[(size_t)k * K + k] - rs + Nsum * th[k];
    }
    for (int i = 0; i < K1; ++i)
        for (int j = 0; j < K1; ++j) H[(size_t)i * K1 + j] = HF[(size_t)i * K + j] + S[(size_t)i * K1 + j];

    /* PD test (np.all(eigvals > 0), stm.py:1017) restated as "Cholesky succeeds" */
    int repair = 0;
    if (chol_lower(K1, H, Lo) != 0) {
        make_pd(K1, H);
        repair = 1;
        if (chol_lower(K1, H, Lo) != 0) {
            for (int i = 0; i < K1; ++i) H[(size_t)i * K1 + i] += 1e-5;
            repair = 2;
        }
    }
    /* decompose_hessian — stm.py:1031-1050 */
    int upper = 0;
    if (chol_lower(K1, H, Lo) != 0) {
        make_pd(K1, H);
        repair += 4;
        if (chol_lower(K1, H, Lo) != 0) {
            make_pd(K1, H);
            for (int i = 0; i < K1; ++i) H[(size_t)i * K1 + i] += 1e-5;
            repair += 8;
            upper = 1; /* scipy.linalg.cholesky returns the UPPER factor (stm.py:1046) */
            if (chol_lower(K1, H, Lo) != 0)
                for (int i = 0; i < K1 * K1; ++i) Lo[i] = NAN;
        }
    }
    out->repair = repair;

    /* lower_bound — stm.py:1068-1101 */
    double det_term = 0.0;
    for (int i = 0; i < K1; ++i) det_term -= log(Lo[(size_t)i * K1 + i]);
    double *d = P.d;
    for (int k = 0; k < K1; ++k) d[k] = eta[k] - mu[k];
    double quad = 0.0;
    for (int j = 0; j < K1; ++j) {
        double t = 0.0;
        for (int i = 0; i < K1; ++i) t += d[i] * S[(size_t)i * K1 + j];
        quad += t * d[j];
    }
    double ll = 0.0;
    for (int v = 0; v < n; ++v) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += th[k] * (B[(size_t)k * n + v] * ex[k]);
        ll += log(s) * c[v];
    }
    out->bound = ll + det_term - 0.5 * quad - sigmaentropy;

    /* optimize_nu — stm.py:1052-1066: nu = inv(triu(L.T)) @ inv(triu(L.T)).T */
    if (!upper) {
        /* Li = inv(L) (lower); U^-1 = Li^T; nu = Li^T Li */
        memset(Li, 0, sizeof(double) * K1 * K1);
        for (int j = 0; j < K1; ++j) {
            Li[(size_t)j * K1 + j] = 1.0 / Lo[(size_t)j * K1 + j];
            for (int i = j + 1; i < K1; ++i) {
                double t = 0.0;
                for (int k = j; k < i; ++k) t += Lo[(size_t)i * K1 + k] * Li[(size_t)k * K1 + j];
                Li[(size_t)i * K1 + j] = -t / Lo[(size_t)i * K1 + i];
            }
        }
        for (int i = 0; i < K1; ++i)
            for (int j = 0; j <= i; ++j) {
                double t = 0.0;
                for (int k = i; k < K1; ++k) t += Li[(size_t)k * K1 + i] * Li[(size_t)k * K1 + j];
                nu_out[(size_t)i * K1 + j] = t;
                nu_out[(size_t)j * K1 + i] = t;
            }
    } else {
        /* L is upper: triu(L.T) keeps only the diagonal */
        memset(nu_out, 0, sizeof(double) * K1 * K1);
        for (int i = 0; i < K1; ++i) {
            double il = 1.0 / Lo[(size_t)i * K1 + i];
            nu_out[(size_t)i * K1 + i] = il * il;
        }
    }

    /* update_z — stm.py:1103-1118: phi = (B o e) * (sqrt(c)/colsum) * sqrt(c) */
    for (int k = 0; k < K; ++k)
        for (int v = 0; v < n; ++v)
            phi_out[(size_t)k * n + v] = ((B[(size_t)k * n + v] * ex[k]) * (sqrt(c[v]) / colsum[v])) * sqrt(c[v]);
}

typedef struct {
    int K, V, nmax;
    int64_t d0, nb, next;
    const int64_t *doc_ptr; const int32_t *word_id; const double *count; const int32_t *aspect;
    const double *beta, *mu, *siginv; double sigmaentropy;
    double *eta, *theta, *work; size_t wsz; double *phi, *nu;
    doc_out_t *outs;
    pthread_mutex_t lock;
} blk_job_t;
typedef struct { blk_job_t *job; int tid; } blk_arg_t;

static void *blk_worker(void *argp) {
    blk_arg_t *arg = (blk_arg_t *)argp;
    blk_job_t *J = arg->job;
    const int K = J->K, K1 = K - 1, V = J->V;
    for (;;) {
        pthread_mutex_lock(&J->lock);
        int64_t j = J->next++;
        pthread_mutex_unlock(&J->lock);
        if (j >= J->nb) break;
        const int64_t d = J->d0 + j;
        const int64_t p0 = J->doc_ptr[d];
        const int n = (int)(J->doc_ptr[d + 1] - p0);
        const int a = J->aspect ? J->aspect[d] : 0;
        infer_doc(K, n, J->word_id + p0, J->count + p0, J->beta + (size_t)a * K * V, V,
                  J->mu + (size_t)d * K1, J->siginv, J->sigmaentropy, J->eta + (size_t)d * K1,
                  J->theta + (size_t)d * K, J->work + J->wsz * (size_t)arg->tid,
                  J->phi + (size_t)K * J->nmax * (size_t)j, J->nu + (size_t)K1 * K1 * (size_t)j,
                  &J->outs[j]);
    }
    return NULL;
}

int stm_oracle_estep(int64_t D, int K, int V, int A,
                     const int64_t *doc_ptr, const int32_t *word_id, const double *count,
                     const int32_t *aspect,
                     const double *beta, const double *mu, const double *siginv, double sigmaentropy,
                     double *eta, double *theta, double *beta_ss, double *sigma_ss, double *bound,
                     double *doc_bound, int32_t *doc_status, int32_t *doc_nit, int32_t *doc_nfev,
                     int32_t *doc_njev, int32_t *doc_repair, int nthreads) {
    if (D < 0 || K < 2 || V < 1 || A < 1) return -1;
    const int K1 = K - 1;
    int nmax = 1;
    for (int64_t d = 0; d < D; ++d) {
        int64_t n = doc_ptr[d + 1] - doc_ptr[d];
        if (n < 0) return -2;
        if (n > nmax) nmax = (int)n;
    }
    for (int64_t i = 0; i < doc_ptr[D]; ++i)
        if (word_id[i] < 0 || word_id[i] >= V) return -3;
    if (aspect)
        for (int64_t d = 0; d < D; ++d)
            if (aspect[d] < 0 || aspect[d] >= A) return -4;

    memset(beta_ss, 0, sizeof(double) * (size_t)A * K * V);
    memset(sigma_ss, 0, sizeof(double) * (size_t)K1 * K1);
    if (nthreads < 1) nthreads = 1;
    if (nthreads > 64) nthreads = 64;
    /* documents are processed in blocks so that accumulation into beta_ss / sigma_ss / bound happens
     * serially in document order (bit-reproducible for any thread count) */
    const int64_t BLK = 256;
    const size_t wsz = doc_work_size(K, nmax);
    double *work = (double *)malloc(sizeof(double) * wsz * (size_t)nthreads);
    double *phi = (double *)malloc(sizeof(double) * (size_t)K * nmax * (size_t)BLK);
    double *nu = (double *)malloc(sizeof(double) * (size_t)K1 * K1 * (size_t)BLK);
    doc_out_t *outs = (doc_out_t *)malloc(sizeof(doc_out_t) * (size_t)BLK);
    if (!work || !phi || !nu || !outs) { free(work); free(phi); free(nu); free(outs); return -5; }
    double total = 0.0;

    for (int64_t d0 = 0; d0 < D; d0 += BLK) {
        const int64_t nb = (D - d0 < BLK) ? (D - d0) : BLK;
        blk_job_t job;
        job.K = K; job.V = V; job.nmax = nmax; job.d0 = d0; job.nb = nb; job.next = 0;
        job.doc_ptr = doc_ptr; job.word_id = word_id; job.count = count; job.aspect = aspect;
        job.beta = beta; job.mu = mu; job.siginv = siginv; job.sigmaentropy = sigmaentropy;
        job.eta = eta; job.theta = theta; job.work = work; job.wsz = wsz; job.phi = phi; job.nu = nu;
        job.outs = outs;
        pthread_mutex_init(&job.lock, NULL);
        int nt = (int)((nb < nthreads) ? nb : nthreads);
        blk_arg_t args[64];
        pthread_t th[64];
        if (nt > 64) nt = 64;
        for (int t = 1; t < nt; ++t) {
            args[t].job = &job; args[t].tid = t;
            pthread_create(&th[t], NULL, blk_worker, &args[t]);
        }
        args[0].job = &job; args[0].tid = 0;
        blk_worker(&args[0]);
        for (int t = 1; t < nt; ++t) pthread_join(th[t], NULL);
        pthread_mutex_destroy(&job.lock);
        for (int64_t j = 0; j < nb; ++j) {
            const int64_t d = d0 + j;
            const int64_t p0 = doc_ptr[d];
            const int n = (int)(doc_ptr[d + 1] - p0);
            const int a = aspect ? aspect[d] : 0;
            double *bs = beta_ss + (size_t)a * K * V;
            const double *ph = phi + (size_t)K * nmax * (size_t)j;
            const double *nv = nu + (size_t)K1 * K1 * (size_t)j;
            for (int i = 0; i < K1 * K1; ++i) sigma_ss[i] += nv[i];
            for (int k = 0; k < K; ++k)
                for (int v = 0; v < n; ++v) bs[(size_t)k * V + word_id[p0 + v]] += ph[(size_t)k * n + v];
            total += outs[j].bound;
            if (doc_bound) doc_bound[d] = outs[j].bound;
            if (doc_status) doc_status[d] = outs[j].status;
            if (doc_nit) doc_nit[d] = outs[j].nit;
            if (doc_nfev) doc_nfev[d] = outs[j].nfev;
            if (doc_njev) doc_njev[d] = outs[j].njev;
            if (doc_repair) doc_repair[d] = outs[j].repair;
        }
    }
    *bound = total;
    free(work); free(phi); free(nu); free(outs);
    return 0;
}

/* ------------------------------------------------------------------------------------------ */
/* exposed pieces                                                                              */
/* ------------------------------------------------------------------------------------------ */

static void setup_doc(doc_t *P, int K, int n, const double *B, const double *c, const double *mu,
                      const double *S, double *buf) {
    P->K = K; P->K1 = K - 1; P->n = n; P->B = B; P->c = c; P->mu = mu; P->S = S;
    double Nsum = 0.0;
    for (int v = 0; v < n; ++v) Nsum += c[v];
    P->Nsum = Nsum; P->Nint = (double)(long long)Nsum;
    double *w = buf;
    P->a = w; w += K; P->et = w; w += K; P->ex = w; w += K; P->d = w; w += K;
    P->w = w; w += (n > K ? n : K); P->xc = w; w += K; P->gc = w; w += K;
    P->have_x = 0; P->f_ok = 0; P->g_ok = 0; P->nfev = 0; P->njev = 0;
    precompute_a(P);
}

double stm_oracle_f(int K, int n, const double *beta_doc, const double *count, const double *mu,
                    const double *siginv, const double *eta) {
    doc_t P;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(7 * K + n + 8));
    setup_doc(&P, K, n, beta_doc, count, mu, siginv, buf);
    double r = obj_f(&P, eta);
    free(buf);
    return r;
}

void stm_oracle_df(int K, int n, const double *beta_doc, const double *count, const double *mu,
                   const double *siginv, const double *eta, double *grad) {
    doc_t P;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(7 * K + n + 8));
    setup_doc(&P, K, n, beta_doc, count, mu, siginv, buf);
    obj_df(&P, eta, grad);
    free(buf);
}

int stm_oracle_bfgs(int K, int n, const double *beta_doc, const double *count, const double *mu,
                    const double *siginv, double *x, double *fun, int *nit, int *nfev, int *njev) {
    doc_t P;
    const int K1 = K - 1;
    double *buf = (double *)malloc(sizeof(double) * (size_t)(7 * K + n + 8));
    double *bw = (double *)malloc(sizeof(double) * (size_t)(6 * K1 + 3 * K1 * K1 + 8));
    setup_doc(&P, K, n, beta_doc, count, mu, siginv, buf);
    int st = bfgs(&P, x, bw, fun, nit);
    *nfev = P.nfev; *njev = P.njev;
    free(buf); free(bw);
    return st;
}
