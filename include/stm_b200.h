/*
 * stm_b200.h — C ABI of the B200-native STM variational-EM core (libstm_b200.so).
 *
 * The reference (mkrcke/strutopy) has no FFI: the hot path sits behind the Python class
 * `STM` (/root/reference/src/modules/stm.py:310).  Each entry point below cites the reference
 * method it replaces; `strutopy_b200/stm.py` is the drop-in `STM` front that binds them through
 * ctypes (INTEGRATION.md shows the stub a maintainer would add to the reference itself).
 *
 * Conventions
 *   - plain C types only; every function returns 0 on success or a negative STM_ERR_* code;
 *     stm_last_error() gives the message.  No exceptions cross the ABI.
 *   - one context per GPU; calls on a context are serialised by the caller.
 *   - "dev" pointers are device pointers OWNED BY THE CALLER (the Python front allocates them as
 *     torch tensors); "host" pointers are host memory (pinned memory makes copies asynchronous).
 *   - device-pointer entry points are asynchronous on `stream` (a cudaStream_t passed as void*);
 *     host-pointer entry points block until their outputs are valid.
 *   - matrices are row-major.  K1 = K-1.  TS = stm_beta_stride(K): word-major beta row stride.
 */
#ifndef STM_B200_H
#define STM_B200_H
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct stm_ctx stm_ctx;

enum {
    STM_OK = 0,
    STM_ERR_INVALID = -1,      /* bad argument (reference: ValueError / AssertionError) */
    STM_ERR_CUDA = -2,         /* CUDA / cuBLAS / cuSOLVER failure */
    STM_ERR_NOT_PD = -3,       /* Sigma not positive definite (reference: stm.py:503-510 raises) */
    STM_ERR_UNSUPPORTED = -4,  /* e.g. non-diagonal siginv, K > 128, document too long for shared memory */
    STM_ERR_NO_CORPUS = -5
};

/* model type / regression mode of update_mu (stm.py:636-711) */
enum {
    STM_MODEL_STM = 0,        /* mode="ols":   LinearRegression (stm.py:690-694) */
    STM_MODEL_CTM = 1,        /* column mean of eta (stm.py:648-651) */
    STM_MODEL_STM_RIDGE = 2,  /* mode="ridge": sklearn Ridge(alpha=0.1) (stm.py:684-688) */
    STM_MODEL_STM_LASSO = 3   /* mode="lasso": sklearn Lasso(alpha=1)   (stm.py:678-682) */
};

/* ---- lifecycle --------------------------------------------------------------------------------- */

/* STM.__init__ (stm.py:311-399) shape part: K topics, V vocabulary, A content-covariate levels
 * (1 when no content model).  Selects `device`, creates cuBLAS/cuSOLVER handles. */
int stm_create(int device, int K, int V, int A, stm_ctx** out);
void stm_destroy(stm_ctx* ctx);
/* message of the last failure on ctx (ctx may be NULL for a failed stm_create) */
const char* stm_last_error(const stm_ctx* ctx);
/* word-major beta row stride in elements: smallest multiple of 4 >= K whose quarter is odd */
int stm_beta_stride(int K);
/* number of this library's own kernel launches issued so far on this context (bench.py's
 * gpu_launches; cuBLAS/cuSOLVER kernels are not counted) */
int64_t stm_launch_count(const stm_ctx* ctx);
/* durations (ms, CUDA events on the stream stm_estep was given) of the two kernel phases of the LAST
 * stm_estep call: ms2[0] = kernel A (per-document BFGS, stm.py:536-545 + SciPy), ms2[1] = kernel B
 * (theta, phi, Hessian, Cholesky pivots, bound, nu; stm.py:546-590).  Blocks until that call is done. */
int stm_estep_kernel_ms(stm_ctx* ctx, double* ms2);
/* tuning of the E-step launch configuration (no reference counterpart; the library reads no environment
 * variables).  Takes effect at the next stm_set_corpus.  Keys: "bfgs_max_warps" (cap on the warps, i.e.
 * documents in flight, per SM of kernel A; occupancy studies). */
int stm_tune(stm_ctx* ctx, const char* key, int value);

/* ---- corpus ------------------------------------------------------------------------------------ */

/* `self.documents` (stm.py:331-332, 366) as CSR, uploaded once per fit: doc_ptr[D+1], word_id[nnz]
 * (unique within a document, 0 <= id < V), count[nnz]; aspect[D] = `beta_index` (stm.py:378, 528)
 * or NULL.  Host pointers.  Builds the document length classes used to size shared memory. */
int stm_set_corpus(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr, const int32_t* word_id,
                   const float* count, const int32_t* aspect);

/* ---- packed sufficient-statistics buffer (the one all-reduced across GPUs) ---------------------
 * fp64, layout for p prevalence covariates:
 *   [0] beta_ss_t  A*V*TS   word-major phi sums        (stm.py:515, 585-588)
 *   [1] sigma_ss   K1*K1    sum of nu                   (stm.py:514, 582)
 *   [2] bound      1        sum of per-document bounds  (stm.py:592)
 *   [3] n_docs     1
 *   [4] sum_eta    K1       [5] sum_x p      [6] xtx p*p      [7] xte p*K1     [8] ete K1*K1
 * offsets[10]: start of each of the 9 segments, offsets[9] = total length (doubles). */
int stm_stats_layout(const stm_ctx* ctx, int p, int64_t* offsets);

/* ---- E-step ------------------------------------------------------------------------------------ */

/* E-step prologue (stm.py:497-501) on device: Cholesky of sigma (cuSOLVER potrf) ->
 * prior_dev[K1+1] = { diag(siginv)[K1], sigmaentropy }.  info_dev (int, device) receives the
 * potrf info (0 = ok, >0 = not PD).  sigma_dev is K1*K1 and is not modified. */
int stm_prologue(stm_ctx* ctx, const double* sigma_dev, double* prior_dev, int* info_dev, void* stream);

/* STM.E_step document loop (stm.py:519-593) — ONE kernel launch per document length class.
 *   beta_t_dev   float [A][V][TS]   word-major beta (padding columns zero)
 *   mu_dev       double [D][K1]
 *   prior_dev    double [K1+1]      from stm_prologue
 *   eta_dev      double [D][K1]     in: warm start (stm.py:539)  out: optimum (stm.py:546)
 *   theta_dev    double [D][K]      out (stm.py:547-549)
 *   stats_dev    packed buffer; segments 0-3 are (re)written
 *   doc_bound_dev double [D]; doc_info_dev int32 [D] (status | nit<<4 | repair<<24);
 *   doc_nfev_dev int32 [D]          per-document diagnostics (always written).  doc_nfev counts the objective
 *                                   evaluations the device made: fewer than SciPy's nfev, because trial points are
 *                                   memoised two deep and a line search whose failure is already decided is not
 *                                   replayed — status, nit and eta are those of the full replay */
int stm_estep(stm_ctx* ctx, const float* beta_t_dev, const double* mu_dev, const double* prior_dev,
              double* eta_dev, double* theta_dev, double* stats_dev, double* doc_bound_dev,
              int32_t* doc_info_dev, int32_t* doc_nfev_dev, void* stream);

/* ---- M-step ------------------------------------------------------------------------------------ */

/* local moments of (eta, X) for update_mu/update_sigma (stm.py:636-728) into stats segments 4-8
 * (cuBLAS).  x_dev: double [D][p] design matrix (already one-hot encoded where the reference
 * would, stm.py:669-671); p may be 0 for CTM. */
int stm_moments(stm_ctx* ctx, const double* eta_dev, const double* x_dev, int p, double* stats_dev,
                void* stream);

/* M_step (stm.py:622-634) from the (all-reduced) statistics:
 *   update_mu   (stm.py:636-711): centred min-norm OLS (sklearn LinearRegression semantics,
 *               cond=1e-6), intercept dropped: gamma_t_dev [p][K1], mu_dev[D][K1] = X gamma'
 *               (CTM: mu = column mean of eta; STM_MODEL_STM_RIDGE / _LASSO: the reference's
 *               regularised modes — centred Ridge(0.1) system, or sklearn's coordinate descent
 *               with its duality-gap stop, both evaluated on the reduced moments)
 *   update_sigma(stm.py:713-728): sigma_dev [K1][K1], shrunk by sigprior
 *   update_beta (stm.py:739-745): beta_t_dev float [A][V][TS] (A>1: the reference normalises over
 *               the topic axis, kept), optionally beta64_t_dev double [A][V][TS] (may be NULL)
 * n_total = global number of documents (stats[3] after the all-reduce). */
int stm_mstep(stm_ctx* ctx, const double* stats_dev, const double* x_dev, int p, int model,
              double sigprior, double* gamma_t_dev, double* mu_dev, double* sigma_dev,
              float* beta_t_dev, double* beta64_t_dev, void* stream);

/* ---- layout helpers (device) ------------------------------------------------------------------- */

/* reference-layout beta double [A][K][V] (device) -> word-major float [A][V][TS] (device) */
int stm_beta_to_wordmajor(stm_ctx* ctx, const double* beta_kv_dev, float* beta_t_dev, void* stream);
/* word-major double [A][V][TS] (device) -> reference layout double [A][K][V] (device) */
int stm_wordmajor_to_kv(stm_ctx* ctx, const double* src_t_dev, double* dst_kv_dev, void* stream);

/* ---- host-buffer entry point: what a reference-side binding calls ------------------------------
 * One full E_step() with the reference's own argument layout, HOST pointers, fp64:
 *   beta [A][K][V], mu [D][K1], siginv [K1][K1] (must be diagonal, as stm.py:501 produces),
 *   sigmaentropy, eta [D][K1] in/out, theta [D][K] out, beta_ss [A][K][V] out, sigma_ss [K1][K1]
 *   out, bound out; doc_bound/doc_status/doc_nit/doc_repair [D] may be NULL.
 * Copies in, runs stm_estep, copies out; blocks until done. */
int stm_estep_host(stm_ctx* ctx, const double* beta, const double* mu, const double* siginv,
                   double sigmaentropy, double* eta, double* theta, double* beta_ss,
                   double* sigma_ss, double* bound, double* doc_bound, int32_t* doc_status,
                   int32_t* doc_nit, int32_t* doc_repair);

/* ---- held-out likelihood, document completion ---------------------------------
 * Replaces eval_heldout(heldout, theta, beta), /root/reference/src/modules/heldout.py:88-97:
 *   doc_ll[d] = sum_w c_w log(theta_d . beta[:, w]) / sum_w c_w   over the held-out words of d,
 *   mean = np.mean(doc_ll)  (an empty held-out document gives NaN, as in the reference).
 * stm_heldout: device pointers, the fit's own buffers (held-out CSR, theta double [D][K], word-major
 * float beta [V][TS]); asynchronous on `stream`; doc_ll_dev [D] and mean_dev [1] are outputs.
 * stm_heldout_host: HOST pointers in the reference's layouts (theta double [D][K], beta double
 * [K][V], kept in fp64 on the device); doc_ll may be NULL; blocks until done.  A = 1 only. */
int stm_heldout(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr_dev, const int32_t* word_id_dev,
                const float* count_dev, const double* theta_dev, const float* beta_t_dev,
                double* doc_ll_dev, double* mean_dev, void* stream);
/* as stm_heldout, with the fit's fp64 master copy of beta (word-major double [V][TS], the beta64_t buffer stm_mstep
 * fills): the reference scores with its float64 beta, and an M-step entry below the fp32 range must not become log(0) */
int stm_heldout64(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr_dev, const int32_t* word_id_dev,
                  const float* count_dev, const double* theta_dev, const double* beta64_t_dev,
                  double* doc_ll_dev, double* mean_dev, void* stream);
int stm_heldout_host(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr, const int32_t* word_id,
                     const float* count, const double* theta, const double* beta_kv, double* doc_ll,
                     double* mean);

/* ---- spectral initialisation of beta ------------------------------------------
 * Replaces spectral_init / gram / fastAnchor / recover_l2, /root/reference/src/modules/stm.py:30-296,
 * over the context's resident corpus (stm_set_corpus).  Two phases, so that a document-sharded fit
 * needs ONE all-reduce (of gram_dev) between them:
 *   keep [n_keep]   HOST: the kept word ids in the reference's order, np.argsort(-wprob)[:maxV]
 *                   (stm.py:57; left to the caller so that ties break exactly as NumPy breaks them)
 *   gram_dev        DEVICE double [n_keep*n_keep + n_keep], caller-owned.  After stm_spectral_gram: this rank's
 *                   part of Htilde'Htilde (stm.py:145-149) as the PACKED row-major upper triangle in the first
 *                   n_keep (n_keep + 1) / 2 entries (index of (i, j >= i): i n - i (i - 1) / 2 + j - i), and
 *                   diag(Hhat) (stm.py:146) in the last n_keep entries — the two slices a document-sharded fit
 *                   all-reduces.  stm_spectral_finish expands it in place and uses it as workspace.
 * stm_spectral_finish: Q = gram - Hhat and the row-sum assertion (stm.py:149-154; failing it returns
 * STM_ERR_INVALID with the reference's message), fastAnchor (stm.py:160-226), recover_l2
 * (stm.py:229-296; the per-word QP is solved exactly as NNLS), beta_new[:, keep] = beta, + 0.001/V,
 * total-sum normalisation (stm.py:78-82).
 *   wprob_keep [n_keep] HOST: wprob[keep] (stm.py:53-59)
 *   beta_kv_dev     DEVICE double [K][V] out (reference layout)
 *   anchor_out [K]  HOST out (may be NULL): anchors as indices into keep (fastAnchor's return value)
 * Both block until done. */
int stm_spectral_gram(stm_ctx* ctx, int n_keep, const int32_t* keep, double* gram_dev, void* stream);
int stm_spectral_finish(stm_ctx* ctx, int n_keep, const int32_t* keep, const double* wprob_keep,
                        double* gram_dev, double* beta_kv_dev, int32_t* anchor_out, void* stream);

/* ---- content-covariate update of beta -------------------------------------------
 * Replaces update_beta with lda_beta=False -> STM.mnreg, /root/reference/src/modules/stm.py:746-853:
 * per word, sklearn PoissonRegressor(fit_intercept=False, alpha) of the (A K) counts beta_ss[a][k][v] on
 * one-hot topic / aspect / interaction covariates (stm.py:769-793), kappa = coefficients (stm.py:841),
 * beta = softmax over the vocabulary of (m + covar @ kappa), split by aspect (stm.py:847-853).
 *   stats_dev    packed statistics (segment 0 = beta_ss, after the all-reduce)
 *   logm_dev     double [V]: m = log(wcounts) - log(sum wcounts) (stm.py:795-797)
 *   alpha        the L2 penalty (the reference uses 250, stm.py:758)
 *   word_column  -1: word v is regressed on its own column v; >= 0: EVERY word is regressed on this
 *                column — the reference as written uses 1 (`counts[:, [1]]`, stm.py:825)
 *   beta_t_dev   float [A][V][TS] out; beta64_t_dev double [A][V][TS] out or NULL
 *   kappa_dev    double [K+A+A*K+1][V] out or NULL (row K is the reference's empty column)
 * The minimiser is unique (strictly convex); it is computed by damped Newton to ~1e-12, sklearn's lbfgs
 * stops at a 1e-5 gradient: agreement ~1e-8.  A >= 2, A <= 8.  Blocks until done.  * kappa_dev, when it is the buffer the previous successful call on this context wrote (same word_column),
 * seeds the Newton iteration: the problem is strictly convex, so the minimiser is the same; it is reached in 1-3
 * steps instead of 5-9. */
int stm_update_kappa(stm_ctx* ctx, const double* stats_dev, const double* logm_dev, double alpha,
                     int word_column, float* beta_t_dev, double* beta64_t_dev, double* kappa_dev, void* stream);

/* ---- synthetic corpus sampler ---------------------------------------------------
 * Replaces CorpusCreation.sample_documents, /root/reference/src/modules/generate_docs.py:293-316:
 * document d ~ Multinomial(n_words, theta_d beta), drawn on the device as n_words (topic, word) pairs per
 * document so that theta @ beta (D x V, generate_docs.py:297) is never materialised.  Philox4x32-10
 * keyed by `seed`: stateless and reproduced bit for bit by oracle/corpus_numpy.py.
 *   theta_dev   double [D][K]  (rows need not be normalised: the inverse CDF is scaled by the row sum)
 *   beta_kv_dev double [K][V]
 *   doc_ptr_dev int64 [D+1], word_id_dev int32 [D*n_words], count_dev float [D*n_words]: CSR out,
 *               ids ascending within a document; *nnz_out (HOST) = doc_ptr[D].  1 <= n_words <= 4096.
 * K and V are the context's.  Blocks until done. */
int stm_sample_corpus(stm_ctx* ctx, int64_t D, int n_words, const double* theta_dev,
                      const double* beta_kv_dev, uint64_t seed, int64_t* doc_ptr_dev, int32_t* word_id_dev,
                      float* count_dev, int64_t* nnz_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif
