/*
 * stm_oracle.h — CPU fp64 restatement of the reference's variational E-step (TEST INFRASTRUCTURE).
 *
 * This is the parity oracle for the CUDA path, NOT a product code path: only tests/,
 * __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may load it.
 *
 * Parity pinning (see oracle/README.md): checked against the live reference
 * (/root/reference/src/modules/stm.py, imported with tools/ref_shims.py) through the committed
 * fixtures the .npz files under tests/golden, incl. the reference's shipped known-answer ELBO
 * src/artifacts/reference_model/{50,70}/lower_bound.pickle[0].
 */
#ifndef STM_ORACLE_H
#define STM_ORACLE_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* E-step prologue — stm.py:497-501. sigma (K1 x K1, row-major) -> siginv (K1 x K1; diagonal by the
 * reference's element-wise product quirk) and sigmaentropy.  Returns 0, or -1 if sigma is not PD. */
int stm_oracle_prologue(int K1, const double *sigma, double *siginv, double *sigmaentropy);

/* One full E-step over D documents — stm.py:519-597 with scipy's BFGS (scipy/optimize/_optimize.py:
 * 1345-1526), DCSRCH (_dcsrch.py) and the Wolfe-2 fallback (_linesearch.py:343-634) restated.
 *
 *  doc_ptr[D+1], word_id[nnz], count[nnz]   CSR bag of words (ids unique within a document)
 *  aspect[D] or NULL                        content-covariate level per document (stm.py:527-532)
 *  beta[A*K*V]                              word-topic matrix/matrices, row-major
 *  mu[D*K1], siginv[K1*K1], sigmaentropy    prior mean per doc, "inverse" covariance, entropy term
 *  eta[D*K1]            in: warm start (stm.py:539)   out: BFGS result (stm.py:546)
 *  theta[D*K]           out (stm.py:547-549)
 *  beta_ss[A*K*V], sigma_ss[K1*K1]          out, zeroed here then accumulated (stm.py:513-515,582-590)
 *  bound                out: sum_d bound_d (stm.py:592)
 *  doc_bound/doc_status/doc_nit/doc_nfev/doc_njev/doc_repair [D] or NULL: per-document diagnostics
 *      status: scipy warnflag (0 ok, 1 maxiter, 2 precision loss, 3 nan)
 *      repair: 0 none, 1 make_pd in hessian(), 2 make_pd + 1e-5; +4 / +8 for the
 *              decompose_hessian fallbacks (stm.py:1041-1048)
 *  nthreads             worker threads (pthreads) over documents (accumulation stays in document order)
 * Returns 0 on success, <0 on invalid arguments.
 */
int stm_oracle_estep(int64_t D, int K, int V, int A,
                     const int64_t *doc_ptr, const int32_t *word_id, const double *count,
                     const int32_t *aspect,
                     const double *beta, const double *mu, const double *siginv, double sigmaentropy,
                     double *eta, double *theta, double *beta_ss, double *sigma_ss, double *bound,
                     double *doc_bound, int32_t *doc_status, int32_t *doc_nit, int32_t *doc_nfev,
                     int32_t *doc_njev, int32_t *doc_repair, int nthreads);

/* Exposed pieces, for unit tests against scipy itself. */

/* objective and (quirky) gradient of one document — stm.py:920-958 */
double stm_oracle_f(int K, int n, const double *beta_doc /*K x n*/, const double *count,
                    const double *mu, const double *siginv, const double *eta);
void stm_oracle_df(int K, int n, const double *beta_doc, const double *count,
                   const double *mu, const double *siginv, const double *eta, double *grad);

/* scipy.optimize.minimize(method="BFGS") on that objective; x in/out. returns warnflag */
int stm_oracle_bfgs(int K, int n, const double *beta_doc, const double *count,
                    const double *mu, const double *siginv, double *x,
                    double *fun, int *nit, int *nfev, int *njev);

/* Self-check of the CUDA kernel's line-search shortcut, the curvature certificate (see stm_oracle.c): returns the
 * counters of the checks made since the last call in out5 = {DCSRCH searches in which the certificate held, of those:
 * a step accepted afterwards, _zoom calls in which it held, of those accepted afterwards, trials the kernel does not
 * make}, then enables / disables the check and zeroes the counters.  Off by default. */
void stm_oracle_shortcut_check(int enable, long long *out5);

#ifdef __cplusplus
}
#endif
#endif
