// stm_b200.cu — host side of libstm_b200.so: context, corpus residency, launch configuration,
// M-step (cuBLAS moments + cuSOLVER factorisations + small fp64 kernels), and the C ABI declared in
// include/stm_b200.h.  Reference behaviour cited per function (stm.py = /root/reference/src/modules/stm.py).
#include "../../include/stm_b200.h"
#include "estep_kernel.cuh"

#include <cublas_v2.h>
#include <cusolverDn.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

namespace {

thread_local std::string g_create_error;

struct LengthClass {
    int n_cap = 0;       // tile rows per warp
    int J = 0;           // words per lane per pass (template parameter)
    int n_docs = 0;
    int* d_docs = nullptr;
    // kernel A (BFGS) / kernel B (post-optimisation): warps per CTA, bytes of shared memory per warp, CTAs
    int warps = 0, smem_per_warp = 0, grid = 0;
    int smem_small = 0, tm_warps = 0, tm_cols = 0;   // kernel A: small block bytes, TMEM-resident warps / columns each
    int post_warps = 0, post_smem_per_warp = 0, post_grid = 0;
    int post_groups = 0;   // > 0: group version of kernel B, this many documents per CTA
    int post_gw = 3;       //      warps per document there
    // host API (stm_estep_host): the class's documents split by document-index chunk, so that the copies of one
    // chunk overlap the kernels of another; chunk_off[c] .. chunk_off[c+1] index d_chunk_docs
    int* d_chunk_docs = nullptr;
    std::vector<int> chunk_off;
};

}  // namespace

struct stm_ctx {
    int device = 0, K = 0, K1 = 0, V = 0, A = 1, TS = 0, KPL = 0;
    int sm_count = 0, max_smem = 0;
    int64_t D = 0, nnz = 0;
    int n_max = 0;
    long long* d_doc_ptr = nullptr;
    int* d_word_id = nullptr;
    float* d_count = nullptr;
    int* d_aspect = nullptr;
    std::vector<LengthClass> classes;
    unsigned int* d_queues = nullptr;
    unsigned long long* d_dbg = nullptr;
    double* d_scratch = nullptr;
    long long scratch_stride = 0;
    int max_warps_total = 0;
    double* d_sigma_rep = nullptr;
    int n_rep = 0;
    // M-step workspaces
    double* d_ones = nullptr;      // [D]
    double* d_msmall = nullptr;    // small fp64 workspace
    int64_t msmall_len = 0;
    double* d_colsum_part = nullptr;
    int colsum_blocks = 0;
    double* d_potrf_work = nullptr;
    int potrf_lwork = 0;
    double* d_syevd_work = nullptr;
    int syevd_lwork = 0;
    int syevd_p = -1;
    int* d_info = nullptr;
    // host-API device buffers
    double *h_beta_kv = nullptr, *h_mu = nullptr, *h_eta = nullptr, *h_theta = nullptr, *h_stats = nullptr,
           *h_prior = nullptr, *h_doc_bound = nullptr, *h_bss_kv = nullptr;
    float* h_beta_t = nullptr;
    int32_t *h_doc_info = nullptr, *h_doc_nfev = nullptr;
    bool host_bufs = false;
    cublasHandle_t cublas = nullptr;
    cusolverDnHandle_t cusolver = nullptr;
    int64_t launches = 0;
    // stm_tune: cap on kernel A's warps per CTA; document chunks of the host API
    int tune_bfgs_max_warps = STM_BFGS_MAX_THREADS / 32;
    int tune_host_chunks = 3;   // r02 A/B at C3, N=1: e2e / value 0.971 (1 chunk), 1.005 (3), 0.999 (4), 0.989 (5)
    double* d_kappa_lin = nullptr;        // stm_update_kappa workspace
    int64_t kappa_lin_len = 0;
    const double* kappa_warm = nullptr;   // stm_update_kappa: coefficient buffer of the last successful call (warm start)
    int kappa_warm_column = -2;
    cudaEvent_t ev[3] = {nullptr, nullptr, nullptr};   // E-step phases: before kernel A, between, after kernel B
    // host API (stm_estep_host): compute stream, copy-in stream, copy-out stream; per-chunk events
    // (inputs landed, kernel A done, kernel B done)
    static constexpr int MAX_CHUNKS = 8;
    int n_chunks = 1;
    std::vector<int64_t> chunk_lo;                     // n_chunks + 1 document bounds
    cudaStream_t host_stream = nullptr, in_stream = nullptr, copy_stream = nullptr;
    cudaEvent_t ev_in[MAX_CHUNKS] = {}, ev_a[MAX_CHUNKS] = {}, ev_b[MAX_CHUNKS] = {}, ev_misc[2] = {};
    std::string err;
};

namespace {

// Every C-ABI entry runs on the context's device and puts the caller's current device back on return (a caller with
// two contexts, or torch's notion of the current device, must not be disturbed).
struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) prev = -1;
        if (prev != dev) ok = (cudaSetDevice(dev) == cudaSuccess);
        else prev = -1;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
    DeviceGuard(const DeviceGuard&) = delete;
    DeviceGuard& operator=(const DeviceGuard&) = delete;
};
#define STM_ON_DEVICE(ctx)                                                                     \
    DeviceGuard dev_guard_((ctx)->device);                                                     \
    if (!dev_guard_.ok) return fail((ctx), STM_ERR_CUDA, "cudaSetDevice failed")

int fail(stm_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}
#define CU(call)                                                                                   \
    do {                                                                                           \
        cudaError_t e_ = (call);                                                                   \
        if (e_ != cudaSuccess)                                                                     \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));    \
    } while (0)
#define CB(call)                                                                                   \
    do {                                                                                           \
        cublasStatus_t s_ = (call);                                                                \
        if (s_ != CUBLAS_STATUS_SUCCESS)                                                           \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": cublas status " + std::to_string((int)s_)); \
    } while (0)
#define CS(call)                                                                                   \
    do {                                                                                           \
        cusolverStatus_t s_ = (call);                                                              \
        if (s_ != CUSOLVER_STATUS_SUCCESS)                                                         \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": cusolver status " + std::to_string((int)s_)); \
    } while (0)

int beta_stride(int K) {
    int ts = (K + 3) / 4 * 4;
    if (((ts / 4) & 1) == 0) ts += 4;
    return ts;
}

// mirror the carve-up at the top of stm::bfgs_kernel: a small block per warp (K-vectors, line-search
// state, mbarrier) and a tile block per warp whose tile lives in shared memory (not TMEM)
size_t bfgs_smem_small(int KPL) {
    const int KVS = KPL * 32 + 8;
    const size_t total = (size_t)4 * KVS * 8 + sizeof(stm::LsState) + 8;
    return (total + 127) & ~(size_t)127;
}
size_t bfgs_smem_tile(int n_cap, int TS, int KPL) {
    const size_t tile = ((size_t)n_cap * TS * 4 + 127) & ~(size_t)127;
    const int KVS = KPL * 32 + 8;
    const size_t total = tile + (size_t)((n_cap + 1) & ~1) * 4 + (n_cap > 3 * KVS ? (size_t)n_cap * 8 : 0);
    return (total + 127) & ~(size_t)127;
}
// mirrors the carve-up at the top of stm::post_kernel
size_t post_smem_per_warp(int n_cap, int TS, int K1, int KPL) {
    const int HS = K1 | 1;
    size_t tile = (size_t)n_cap * TS * 4;
    const size_t hb = (size_t)K1 * HS * 8;
    if (hb > tile) tile = hb;
    tile = (tile + 127) & ~(size_t)127;
    const int KVS = KPL * 32 + 8;
    const size_t total = tile + (size_t)n_cap * 16 + (size_t)4 * KVS * 8 + (size_t)n_cap * 4 +
                         (size_t)((n_cap + 1) & ~1) * 4 + 8;
    return (total + 127) & ~(size_t)127;
}

// mirrors the carve-up at the top of stm::post_group_kernel
size_t post_group_smem(int n_cap, int TS, int K1, int KPL, int gw) {
    const int HS = K1 | 1;
    size_t tile = (size_t)(n_cap + 1) * TS * 4;
    const size_t hb = (size_t)K1 * HS * 8;
    if (hb > tile) tile = hb;
    tile = (tile + 127) & ~(size_t)127;
    size_t wb = (size_t)(n_cap + 4) * 16;
    if (wb < (size_t)4 * stm::post_ust(gw) * 8) wb = (size_t)4 * stm::post_ust(gw) * 8;
    const int KVS = KPL * 32 + 8;
    const size_t total = tile + wb + (size_t)4 * KVS * 8 + (size_t)stm::POST_RED * 8 + (size_t)n_cap * 4 +
                         (size_t)((n_cap + 1) & ~1) * 4 + 16;
    return (total + 127) & ~(size_t)127;
}
// warps per document of the group version of kernel B (one 4x4 patch of the lower triangle per thread)
int post_group_warps(int K1, int KPL) {
    if (K1 <= 52) return 3;
    if (KPL == 2) return 5;
    if (KPL == 3) return 10;
    return K1 <= 100 ? 11 : 17;
}

}  // namespace
// one translation unit per KPL (estep_inst.cu compiled with -DSTM_KPL=n) so the 16 kernel
// instantiations build in parallel
cudaError_t stm_launch_bfgs_kpl1(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_bfgs_kpl2(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_bfgs_kpl3(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_bfgs_kpl4(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_kpl1(const stm::EstepParams&, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_kpl2(const stm::EstepParams&, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_kpl3(const stm::EstepParams&, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_kpl4(const stm::EstepParams&, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_group_kpl1(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_group_kpl2(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_group_kpl3(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
cudaError_t stm_launch_post_group_kpl4(const stm::EstepParams&, int, int, int, size_t, cudaStream_t);
namespace {

cudaError_t launch_bfgs(int KPL, const stm::EstepParams& P, int J, int grid, int block, size_t smem,
                        cudaStream_t st) {
    switch (KPL) {
        case 1: return stm_launch_bfgs_kpl1(P, J, grid, block, smem, st);
        case 2: return stm_launch_bfgs_kpl2(P, J, grid, block, smem, st);
        case 3: return stm_launch_bfgs_kpl3(P, J, grid, block, smem, st);
        default: return stm_launch_bfgs_kpl4(P, J, grid, block, smem, st);
    }
}
cudaError_t launch_post(int KPL, const stm::EstepParams& P, int grid, int block, size_t smem, cudaStream_t st) {
    switch (KPL) {
        case 1: return stm_launch_post_kpl1(P, grid, block, smem, st);
        case 2: return stm_launch_post_kpl2(P, grid, block, smem, st);
        case 3: return stm_launch_post_kpl3(P, grid, block, smem, st);
        default: return stm_launch_post_kpl4(P, grid, block, smem, st);
    }
}

// ---- small device kernels ----------------------------------------------------------------------

// E-step epilogue: sigma_ss = mirror(sum of replicas), bound = ordered sum of doc_bound, n_docs.
__global__ void estep_epilogue_kernel(const double* __restrict__ sig_rep, int n_rep, int K1,
                                      const double* __restrict__ doc_bound, long long D,
                                      double* __restrict__ sigma_ss, double* __restrict__ bound,
                                      double* __restrict__ ndocs) {
    __shared__ double red[1024];
    const int t = threadIdx.x;
    for (int idx = t; idx < K1 * K1; idx += blockDim.x) {
        const int i = idx / K1, j = idx % K1;
        const int lo = (i >= j) ? (i * K1 + j) : (j * K1 + i);
        double s = 0.0;
        for (int r = 0; r < n_rep; ++r) s += sig_rep[(size_t)r * K1 * K1 + lo];
        sigma_ss[idx] = s;
    }
    // fixed-order (deterministic) sum of the per-document bounds
    double s = 0.0;
    for (long long d = t; d < D; d += blockDim.x) s += doc_bound[d];
    red[t] = s;
    __syncthreads();
    for (int o = blockDim.x / 2; o > 0; o >>= 1) {
        if (t < o) red[t] += red[t + o];
        __syncthreads();
    }
    if (t == 0) { *bound = red[0]; *ndocs = (double)D; }
}

// prior = { 1 / L_ii^2, sum log L_ii } from the potrf factor (column-major lower == row-major upper)
__global__ void prior_from_chol_kernel(const double* __restrict__ L, int K1, const int* __restrict__ info,
                                       double* __restrict__ prior) {
    if (threadIdx.x == 0) {
        double ent = 0.0;
        const bool bad = (*info != 0);
        for (int i = 0; i < K1; ++i) {
            const double l = L[(size_t)i * K1 + i];
            ent += log(l);
            const double il = 1.0 / l;
            prior[i] = bad ? nan("") : il * il;
        }
        prior[K1] = bad ? nan("") : ent;
    }
}

__global__ void beta_to_wordmajor_kernel(const double* __restrict__ src, float* __restrict__ dst, int A,
                                         int K, int V, int TS) {
    const long long total = (long long)A * V * TS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % TS);
        const long long av = i / TS;
        const int v = (int)(av % V);
        const int a = (int)(av / V);
        dst[i] = (k < K) ? (float)src[((size_t)a * K + k) * V + v] : 0.0f;
    }
}
__global__ void wordmajor_to_kv_kernel(const double* __restrict__ src, double* __restrict__ dst, int A,
                                       int K, int V, int TS) {
    const long long total = (long long)A * K * V;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int v = (int)(i % V);
        const long long ak = i / V;
        const int k = (int)(ak % K);
        const int a = (int)(ak / K);
        dst[i] = src[((size_t)a * V + v) * TS + k];
    }
}

// update_beta, A == 1: column sums over the vocabulary, two-stage and order-deterministic
__global__ void colsum_partial_kernel(const double* __restrict__ ss, int V, int TS, int rows_per_block,
                                      double* __restrict__ part) {
    const int k = threadIdx.x;
    if (k >= TS) return;
    const int v0 = blockIdx.x * rows_per_block;
    const int v1 = min(V, v0 + rows_per_block);
    double s = 0.0;
    for (int v = v0; v < v1; ++v) s += ss[(size_t)v * TS + k];
    part[(size_t)blockIdx.x * TS + k] = s;
}
__global__ void colsum_final_kernel(const double* __restrict__ part, int nblk, int TS, double* __restrict__ out) {
    const int k = threadIdx.x;
    if (k >= TS) return;
    double s = 0.0;
    for (int b = 0; b < nblk; ++b) s += part[(size_t)b * TS + k];
    out[k] = s;
}
// beta = beta_ss / rowsum (zero-safe), stm.py:741-745
__global__ void beta_normalise_kernel(const double* __restrict__ ss, const double* __restrict__ rowsum, int V,
                                      int K, int TS, float* __restrict__ beta_t, double* __restrict__ beta64_t) {
    const long long total = (long long)V * TS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % TS);
        double b = 0.0;
        if (k < K) {
            const double rs = rowsum[k];
            if (rs != 0.0) b = ss[i] / rs;
        }
        beta_t[i] = (float)b;
        if (beta64_t) beta64_t[i] = b;
    }
}
// A > 1: the reference's np.sum(beta_ss, axis=1) on an A x K x V array sums over TOPICS (stm.py:741)
__global__ void beta_normalise_aspect_kernel(const double* __restrict__ ss, long long AV, int K, int TS,
                                             float* __restrict__ beta_t, double* __restrict__ beta64_t) {
    for (long long r = blockIdx.x * (long long)blockDim.x + threadIdx.x; r < AV;
         r += (long long)gridDim.x * blockDim.x) {
        const double* row = ss + (size_t)r * TS;
        double s = 0.0;
        for (int k = 0; k < K; ++k) s += row[k];
        for (int k = 0; k < TS; ++k) {
            double b = 0.0;
            if (k < K && s != 0.0) b = row[k] / s;
            beta_t[(size_t)r * TS + k] = (float)b;
            if (beta64_t) beta64_t[(size_t)r * TS + k] = b;
        }
    }
}

// update_mu (OLS) + update_sigma on the reduced moments: one block, fp64.
// Workspace layout `w` (doubles): G[p*p] (in: centred Gram, out of syevd: eigenvectors, column-major),
// lam[p], R[p*K1] centred X'eta, gamma_t [p*K1] (out).
__global__ void center_moments_kernel(const double* __restrict__ stats, const long long* __restrict__ off,
                                      int p, int K1, double* __restrict__ G, double* __restrict__ R) {
    const double N = stats[off[3]];
    const double* sum_eta = stats + off[4];
    const double* sum_x = stats + off[5];
    const double* xtx = stats + off[6];
    const double* xte = stats + off[7];
    for (int idx = threadIdx.x; idx < p * p; idx += blockDim.x) {
        const int i = idx / p, j = idx % p;
        G[idx] = xtx[idx] - (sum_x[i] / N) * sum_x[j];
    }
    for (int idx = threadIdx.x; idx < p * K1; idx += blockDim.x) {
        const int i = idx / K1, k = idx % K1;
        R[idx] = xte[idx] - (sum_x[i] / N) * sum_eta[k];
    }
}
// gamma_t [p][K1] = pinv(G) R with scipy.linalg.lstsq(cond=1e-6) semantics on the centred design:
// singular values of Xc are sqrt(lam); those <= 1e-6 * max are dropped (min-norm solution).
// ridge > 0: sklearn Ridge(alpha=ridge) on the centred design, (Xc'Xc + ridge I) gamma = Xc'eta_c
// (stm.py:684-688; _ridge.py's dense 'cholesky' solver solves exactly this system).
__global__ void solve_gamma_kernel(const double* __restrict__ Vec /*col-major p x p*/, const double* __restrict__ lam,
                                   const double* __restrict__ R, int p, int K1, double ridge,
                                   double* __restrict__ tmp /*p*K1*/, double* __restrict__ gamma_t) {
    double lmax = 0.0;
    for (int i = 0; i < p; ++i) lmax = fmax(lmax, lam[i]);
    const double cut = 1e-12 * lmax;  // (1e-6)^2 on eigenvalues of Xc'Xc
    // tmp[e][k] = (V[:,e] . R[:,k]) / lam_e
    for (int idx = threadIdx.x; idx < p * K1; idx += blockDim.x) {
        const int e = idx / K1, k = idx % K1;
        double s = 0.0;
        for (int i = 0; i < p; ++i) s += Vec[(size_t)e * p + i] * R[(size_t)i * K1 + k];
        if (ridge > 0.0) tmp[idx] = s / (fmax(lam[e], 0.0) + ridge);
        else tmp[idx] = (lam[e] > cut && lam[e] > 0.0) ? s / lam[e] : 0.0;
    }
    __syncthreads();
    for (int idx = threadIdx.x; idx < p * K1; idx += blockDim.x) {
        const int i = idx / K1, k = idx % K1;
        double s = 0.0;
        for (int e = 0; e < p; ++e) s += Vec[(size_t)e * p + i] * tmp[(size_t)e * K1 + k];
        gamma_t[idx] = s;
    }
}
// sklearn Lasso(alpha=1, fit_intercept=True).fit(X, eta).coef_ (stm.py:678-682) on the reduced moments.
// sklearn fits each target separately by cyclic coordinate descent (linear_model/_cd_fast.pyx,
// enet_coordinate_descent, installed 1.9.0: gap-safe screening, duality-gap stop at tol * y'y,
// tol = 1e-4, max_iter = 1000) on the centred data with l1_reg = alpha * n_samples.  Every quantity it
// forms is a function of G = Xc'Xc, b = Xc'y_c and yy = y_c'y_c:
//   X_j'R = b_j - (G w)_j,   R'R = yy - 2 w'b + w'G w,   R'y = yy - w'b
// so the same iteration runs here from the all-reduced moments, one thread per topic.
// ws: 3 p doubles per topic (w | XtA | flags).
__global__ void lasso_gamma_kernel(const double* __restrict__ G, const double* __restrict__ B /*p x K1*/,
                                   const double* __restrict__ stats, const long long* __restrict__ off, int p, int K1,
                                   double alpha_per_sample, double* __restrict__ ws, double* __restrict__ gamma_t) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K1) return;
    const double N = stats[off[3]];
    const double se = stats[off[4] + k];
    const double yy = stats[off[8] + (size_t)k * K1 + k] - (se / N) * se;
    const double alpha = alpha_per_sample * N;
    double* w = ws + (size_t)k * 3 * p;
    double* XtA = w + p;
    double* excl = XtA + p;          // 0 active, 1 excluded
    for (int j = 0; j < p; ++j) { w[j] = 0.0; excl[j] = 0.0; }
    const double tol = 1e-4 * yy;
    double gap = 0.0, dual_norm = 0.0;
    auto bj = [&](int j) { return B[(size_t)j * K1 + k]; };
    auto gw = [&](int j) { double s = 0.0; for (int i = 0; i < p; ++i) s += G[(size_t)j * p + i] * w[i]; return s; };
    auto gap_enet = [&]() {
        double wb = 0.0, wgw = 0.0, l1 = 0.0;
        dual_norm = 0.0;
        for (int j = 0; j < p; ++j) {
            const double g = gw(j);
            XtA[j] = bj(j) - g;
            dual_norm = fmax(dual_norm, fabs(XtA[j]));
            wb += w[j] * bj(j);
            wgw += w[j] * g;
            l1 += fabs(w[j]);
        }
        const double R2 = yy - 2.0 * wb + wgw, Ry = yy - wb;
        const double primal = 0.5 * R2 + alpha * l1;
        const double scale = (dual_norm > alpha) ? alpha / dual_norm : 1.0;
        const double dual = -0.5 * (scale * scale) * R2 + scale * Ry;
        gap = primal - dual;
    };
    auto screen = [&](bool first) {
        for (int j = 0; j < p; ++j) {
            if (first) {
                if (G[(size_t)j * p + j] == 0.0) { w[j] = 0.0; excl[j] = 1.0; continue; }
            } else if (excl[j] != 0.0) continue;
            const double xt = XtA[j] / fmax(alpha, dual_norm);
            const double dj = (1.0 - fabs(xt)) / sqrt(G[(size_t)j * p + j]);
            if (dj <= sqrt(2.0 * gap) / alpha) excl[j] = 0.0;
            else { w[j] = 0.0; excl[j] = 1.0; }
        }
    };
    gap_enet();
    if (!(gap <= tol)) {
        screen(true);
        for (int it = 0; it < 1000; ++it) {
            double w_max = 0.0, d_w_max = 0.0;
            for (int j = 0; j < p; ++j) {
                if (excl[j] != 0.0) continue;
                const double gjj = G[(size_t)j * p + j];
                if (gjj == 0.0) continue;
                const double wj = w[j];
                const double tmp = (bj(j) - gw(j)) + wj * gjj;
                const double sgn = (tmp > 0.0) ? 1.0 : ((tmp < 0.0) ? -1.0 : 0.0);
                w[j] = sgn * fmax(fabs(tmp) - alpha, 0.0) / gjj;
                d_w_max = fmax(d_w_max, fabs(w[j] - wj));
                w_max = fmax(w_max, fabs(w[j]));
            }
            if (w_max == 0.0 || d_w_max / w_max <= 1e-4 || it == 999) {
                gap_enet();
                if (gap <= tol) break;
                screen(false);
            }
        }
    }
    for (int j = 0; j < p; ++j) gamma_t[(size_t)j * K1 + k] = w[j];
}
// sigma = ((eta-mu)'(eta-mu) + sigma_ss)/N with the residual Gram expanded from the global moments,
// then the sigprior shrinkage (stm.py:723-728).  mode: 0 STM (mu = X gamma'), 1 CTM (mu = mean eta).
__global__ void sigma_update_kernel(const double* __restrict__ stats, const long long* __restrict__ off, int p,
                                    int K1, int model, const double* __restrict__ gamma_t, double sigprior,
                                    double* __restrict__ tmp /* p*K1 */, double* __restrict__ sigma) {
    const double N = stats[off[3]];
    const double* sigma_ss = stats + off[1];
    const double* sum_eta = stats + off[4];
    const double* xtx = stats + off[6];
    const double* xte = stats + off[7];
    const double* ete = stats + off[8];
    if (model == STM_MODEL_STM) {
        // tmp[i][k] = sum_j xtx[i][j] gamma_t[j][k]
        for (int idx = threadIdx.x; idx < p * K1; idx += blockDim.x) {
            const int i = idx / K1, k = idx % K1;
            double s = 0.0;
            for (int j = 0; j < p; ++j) s += xtx[(size_t)i * p + j] * gamma_t[(size_t)j * K1 + k];
            tmp[idx] = s;
        }
        __syncthreads();
    }
    for (int idx = threadIdx.x; idx < K1 * K1; idx += blockDim.x) {
        const int a = idx / K1, b = idx % K1;
        double cov;
        if (model == STM_MODEL_STM) {
            // (eta - X g')'(eta - X g') = ete - T' - T + g (xtx) g',  T[a][b] = sum_i gamma_t[i][a] xte[i][b]
            double t_ab = 0.0, t_ba = 0.0, q = 0.0;
            for (int i = 0; i < p; ++i) {
                t_ab += gamma_t[(size_t)i * K1 + a] * xte[(size_t)i * K1 + b];
                t_ba += gamma_t[(size_t)i * K1 + b] * xte[(size_t)i * K1 + a];
                q += gamma_t[(size_t)i * K1 + a] * tmp[(size_t)i * K1 + b];
            }
            cov = ete[idx] - t_ab - t_ba + q;
        } else {
            cov = ete[idx] - (sum_eta[a] / N) * sum_eta[b];
        }
        const double s = (cov + sigma_ss[idx]) / N;
        sigma[idx] = (a == b) ? (s * sigprior + (1.0 - sigprior) * s) : ((1.0 - sigprior) * s);
    }
}
__global__ void mu_ctm_kernel(const double* __restrict__ stats, const long long* __restrict__ off, int K1,
                              long long D, double* __restrict__ mu) {
    const double N = stats[off[3]];
    const double* sum_eta = stats + off[4];
    const long long total = D * K1;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        mu[i] = sum_eta[i % K1] / N;
}
__global__ void fill_kernel(double* p, long long n, double v) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
        p[i] = v;
}
__global__ void unpack_info_kernel(const int* __restrict__ info, long long D, int* status, int* nit, int* repair) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < D; i += (long long)gridDim.x * blockDim.x) {
        const int w = info[i];
        if (status) status[i] = w & 0xf;
        if (nit) nit[i] = (w >> 4) & 0xfffff;
        if (repair) repair[i] = (w >> 24) & 0xff;
    }
}

void layout(const stm_ctx* c, int p, int64_t* off) {
    const int64_t K1 = c->K1;
    int64_t o = 0;
    off[0] = o; o += (int64_t)c->A * c->V * c->TS;
    off[1] = o; o += K1 * K1;
    off[2] = o; o += 1;
    off[3] = o; o += 1;
    off[4] = o; o += K1;
    off[5] = o; o += p;
    off[6] = o; o += (int64_t)p * p;
    off[7] = o; o += (int64_t)p * K1;
    off[8] = o; o += K1 * K1;
    off[9] = o;
}

void free_corpus(stm_ctx* c) {
    cudaFree(c->d_doc_ptr); cudaFree(c->d_word_id); cudaFree(c->d_count); cudaFree(c->d_aspect);
    c->d_doc_ptr = nullptr; c->d_word_id = nullptr; c->d_count = nullptr; c->d_aspect = nullptr;
    for (auto& lc : c->classes) { cudaFree(lc.d_docs); cudaFree(lc.d_chunk_docs); }
    c->classes.clear();
    cudaFree(c->d_queues); c->d_queues = nullptr;
    cudaFree(c->d_dbg); c->d_dbg = nullptr;
    cudaFree(c->d_scratch); c->d_scratch = nullptr;
    cudaFree(c->d_sigma_rep); c->d_sigma_rep = nullptr;
    cudaFree(c->d_ones); c->d_ones = nullptr;
    cudaFree(c->d_colsum_part); c->d_colsum_part = nullptr;
    if (c->host_bufs) {
        cudaFree(c->h_beta_kv); cudaFree(c->h_mu); cudaFree(c->h_eta); cudaFree(c->h_theta);
        cudaFree(c->h_stats); cudaFree(c->h_prior); cudaFree(c->h_doc_bound); cudaFree(c->h_bss_kv);
        cudaFree(c->h_beta_t); cudaFree(c->h_doc_info); cudaFree(c->h_doc_nfev);
        c->host_bufs = false;
    }
}

}  // namespace

// =================================================================================================
// C ABI
// =================================================================================================
// ---- held-out likelihood (document completion), /root/reference/src/modules/heldout.py:88-97 ------------
// doc_ll[d] = sum_w c_w log(theta_d . beta[:, w]) / sum_w c_w over the held-out words of document d; one warp per
// document, lane <-> word, theta_d broadcast from shared memory, beta word-major (one contiguous row per word).
template <typename BT>
__global__ void heldout_kernel(const long long* __restrict__ doc_ptr, const int* __restrict__ word_id,
                               const float* __restrict__ count, const double* __restrict__ theta,
                               const BT* __restrict__ beta_t, int K, int TS, long long D, double* __restrict__ doc_ll) {
    extern __shared__ double th_s[];   // [warps][K]
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* th = th_s + (size_t)warp * K;
    for (long long d = (long long)blockIdx.x * nw + warp; d < D; d += (long long)gridDim.x * nw) {
        __syncwarp();
        for (int k = lane; k < K; k += 32) th[k] = theta[(size_t)d * K + k];
        __syncwarp();
        const long long p0 = doc_ptr[d], p1 = doc_ptr[d + 1];
        double ll = 0.0, nsum = 0.0;
        for (long long p = p0 + lane; p < p1; p += 32) {
            const BT* row = beta_t + (size_t)word_id[p] * TS;
            double s = 0.0;
            for (int k = 0; k < K; ++k) s = fma(th[k], (double)row[k], s);
            const double c = (double)count[p];
            ll += c * log(s);
            nsum += c;
        }
        for (int o = 16; o > 0; o >>= 1) {
            ll += __shfl_xor_sync(0xffffffffu, ll, o);
            nsum += __shfl_xor_sync(0xffffffffu, nsum, o);
        }
        if (lane == 0) doc_ll[d] = ll / nsum;   // an empty document gives 0/0 = NaN, like the reference
    }
}
// np.mean(doc_ll): fixed-order tree sum in one block (deterministic)
__global__ void mean_kernel(const double* __restrict__ x, long long n, double* __restrict__ out) {
    __shared__ double sh[1024];
    double a = 0.0;
    for (long long i = threadIdx.x; i < n; i += blockDim.x) a += x[i];
    sh[threadIdx.x] = a;
    __syncthreads();
    for (int o = blockDim.x >> 1; o > 0; o >>= 1) {
        if ((int)threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) *out = sh[0] / (double)n;
}
__global__ void beta_kv_to_wordmajor_f64_kernel(const double* __restrict__ src, double* __restrict__ dst, int K, int V,
                                                int TS) {
    const long long total = (long long)V * TS;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int k = (int)(i % TS);
        const long long v = i / TS;
        dst[i] = (k < K) ? src[(size_t)k * V + v] : 0.0;
    }
}

template <typename BT>
int heldout_launch(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr_dev, const int32_t* word_id_dev,
                   const float* count_dev, const double* theta_dev, const BT* beta_t_dev, double* doc_ll_dev,
                   double* mean_dev, cudaStream_t st) {
    const int warps = 8;
    const int grid = (int)std::min<int64_t>((D + warps - 1) / warps, (int64_t)ctx->sm_count * 8);
    heldout_kernel<BT><<<std::max(grid, 1), warps * 32, sizeof(double) * warps * ctx->K, st>>>(
        reinterpret_cast<const long long*>(doc_ptr_dev), word_id_dev, count_dev, theta_dev, beta_t_dev, ctx->K, ctx->TS,
        (long long)D, doc_ll_dev);
    mean_kernel<<<1, 1024, 0, st>>>(doc_ll_dev, (long long)D, mean_dev);
    ctx->launches += 2;
    CU(cudaGetLastError());
    return STM_OK;
}

extern "C" {

int stm_beta_stride(int K) { return beta_stride(K); }

const char* stm_last_error(const stm_ctx* ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }

int64_t stm_launch_count(const stm_ctx* ctx) { return ctx ? ctx->launches : 0; }

int stm_estep_kernel_ms(stm_ctx* ctx, double* ms2) {
    if (!ctx || !ms2) return STM_ERR_INVALID;
    STM_ON_DEVICE(ctx);
    CU(cudaEventSynchronize(ctx->ev[2]));
    float a = 0.f, b = 0.f;
    CU(cudaEventElapsedTime(&a, ctx->ev[0], ctx->ev[1]));
    CU(cudaEventElapsedTime(&b, ctx->ev[1], ctx->ev[2]));
    ms2[0] = a; ms2[1] = b;
    return STM_OK;
}

// Tuning interface (explicit; the library reads no environment variables).  Takes effect at the next
// stm_set_corpus.  Keys: "bfgs_max_warps" (cap on kernel A's warps = documents in flight per SM; occupancy studies),
// "host_chunks" (1..8: document chunks whose copies stm_estep_host overlaps with the kernels of other chunks).
int stm_tune(stm_ctx* ctx, const char* key, int value) {
    if (!ctx || !key) return STM_ERR_INVALID;
    const std::string k(key);
    if (k == "bfgs_max_warps") {
        if (value < 1) return fail(ctx, STM_ERR_INVALID, "bfgs_max_warps must be >= 1");
        ctx->tune_bfgs_max_warps = value;
        return STM_OK;
    }
    if (k == "host_chunks") {
        if (value < 1 || value > stm_ctx::MAX_CHUNKS) return fail(ctx, STM_ERR_INVALID, "host_chunks must be in 1..8");
        ctx->tune_host_chunks = value;
        return STM_OK;
    }
    return fail(ctx, STM_ERR_INVALID, "stm_tune: unknown key " + k);
}

int stm_create(int device, int K, int V, int A, stm_ctx** out) {
    stm_ctx* ctx = nullptr;
    if (!out) return fail(ctx, STM_ERR_INVALID, "out is NULL");
    *out = nullptr;
    if (K < 2) return fail(ctx, STM_ERR_INVALID, "Number of topics must be >= 2");  // stm.py:393-394
    if (K > 128) return fail(ctx, STM_ERR_UNSUPPORTED, "K > 128 is not supported by the warp-per-document kernel");
    if (V < 1 || A < 1) return fail(ctx, STM_ERR_INVALID, "V and A must be >= 1");
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0)
        return fail(ctx, STM_ERR_CUDA, "no CUDA device available (this library has no CPU fallback)");
    if (device < 0 || device >= ndev) return fail(ctx, STM_ERR_INVALID, "bad device index");
    DeviceGuard guard(device);
    if (!guard.ok) return fail(ctx, STM_ERR_CUDA, "cudaSetDevice failed");
    cudaError_t e;
    cudaDeviceProp prop;
    e = cudaGetDeviceProperties(&prop, device);
    if (e != cudaSuccess) return fail(ctx, STM_ERR_CUDA, cudaGetErrorString(e));
    if (prop.major < 10)
        return fail(ctx, STM_ERR_UNSUPPORTED, "libstm_b200 is built for sm_100a (Blackwell B200) only");
    stm_ctx* c = new stm_ctx();
    c->device = device; c->K = K; c->K1 = K - 1; c->V = V; c->A = A;
    c->TS = beta_stride(K);
    c->KPL = (K + 31) / 32;
    c->sm_count = prop.multiProcessorCount;
    c->max_smem = (int)prop.sharedMemPerBlockOptin;
    if (cublasCreate(&c->cublas) != CUBLAS_STATUS_SUCCESS || cusolverDnCreate(&c->cusolver) != CUSOLVER_STATUS_SUCCESS) {
        delete c;
        return fail(ctx, STM_ERR_CUDA, "cuBLAS / cuSOLVER handle creation failed");
    }
    cublasSetPointerMode(c->cublas, CUBLAS_POINTER_MODE_HOST);
    // small M-step workspace
    c->msmall_len = 4096 + 8LL * c->K1 * c->K1;
    if (cudaMalloc(&c->d_msmall, sizeof(double) * c->msmall_len) != cudaSuccess ||
        cudaMalloc(&c->d_info, sizeof(int) * 4) != cudaSuccess) {
        delete c;
        return fail(ctx, STM_ERR_CUDA, "workspace allocation failed");
    }
    cusolverDnDpotrf_bufferSize(c->cusolver, CUBLAS_FILL_MODE_LOWER, c->K1, c->d_msmall, c->K1, &c->potrf_lwork);
    if (c->potrf_lwork < 1) c->potrf_lwork = 1;
    cudaMalloc(&c->d_potrf_work, sizeof(double) * c->potrf_lwork);
    for (int i = 0; i < 3; ++i) cudaEventCreate(&c->ev[i]);
    cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking);
    cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking);
    for (int i = 0; i < stm_ctx::MAX_CHUNKS; ++i) {
        cudaEventCreateWithFlags(&c->ev_in[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_a[i], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_b[i], cudaEventDisableTiming);
    }
    for (int i = 0; i < 2; ++i) cudaEventCreateWithFlags(&c->ev_misc[i], cudaEventDisableTiming);
    *out = c;
    return STM_OK;
}

void stm_destroy(stm_ctx* c) {
    if (!c) return;
    DeviceGuard guard(c->device);
    free_corpus(c);
    cudaFree(c->d_msmall); cudaFree(c->d_info); cudaFree(c->d_potrf_work); cudaFree(c->d_syevd_work);
    cudaFree(c->d_kappa_lin);
    if (c->cublas) cublasDestroy(c->cublas);
    if (c->cusolver) cusolverDnDestroy(c->cusolver);
    for (int i = 0; i < 3; ++i) if (c->ev[i]) cudaEventDestroy(c->ev[i]);
    if (c->copy_stream) cudaStreamDestroy(c->copy_stream);
    if (c->host_stream) cudaStreamDestroy(c->host_stream);
    if (c->in_stream) cudaStreamDestroy(c->in_stream);
    for (int i = 0; i < stm_ctx::MAX_CHUNKS; ++i) {
        if (c->ev_in[i]) cudaEventDestroy(c->ev_in[i]);
        if (c->ev_a[i]) cudaEventDestroy(c->ev_a[i]);
        if (c->ev_b[i]) cudaEventDestroy(c->ev_b[i]);
    }
    for (int i = 0; i < 2; ++i) if (c->ev_misc[i]) cudaEventDestroy(c->ev_misc[i]);
    delete c;
}

int stm_set_corpus(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr, const int32_t* word_id, const float* count,
                   const int32_t* aspect) {
    if (!ctx) return STM_ERR_INVALID;
    if (D < 0 || !doc_ptr) return fail(ctx, STM_ERR_INVALID, "documents must be specified to establish input space");
    STM_ON_DEVICE(ctx);
    if (doc_ptr[0] != 0) return fail(ctx, STM_ERR_INVALID, "doc_ptr[0] must be 0");
    const int64_t nnz = doc_ptr[D];
    int n_max = 0;
    for (int64_t d = 0; d < D; ++d) {
        const int64_t n = doc_ptr[d + 1] - doc_ptr[d];
        if (n < 0 || n > (1 << 24)) return fail(ctx, STM_ERR_INVALID, "doc_ptr must be non-decreasing");
        n_max = std::max<int>(n_max, (int)n);
    }
    for (int64_t i = 0; i < nnz; ++i)
        if (word_id[i] < 0 || word_id[i] >= ctx->V) return fail(ctx, STM_ERR_INVALID, "word id out of range [0, V)");
    if (aspect) {
        for (int64_t d = 0; d < D; ++d)
            if (aspect[d] < 0 || aspect[d] >= ctx->A) return fail(ctx, STM_ERR_INVALID, "aspect out of range [0, A)");
    }
    // ---- length classes: tile capacity per warp -> warps per CTA -------------------------------
    // (all validation happens BEFORE the previous corpus is released: a failing call leaves the context as it was)
    const int caps[] = {64, 128, 160, 256, 384, 512, 768, 1024, 1536, 2048, 3072, 4096, 6144, 8192};
    const int ncaps = (int)(sizeof(caps) / sizeof(caps[0]));
    std::vector<std::vector<int>> members(ncaps);
    for (int64_t d = 0; d < D; ++d) {
        const int n = (int)(doc_ptr[d + 1] - doc_ptr[d]);
        int ci = 0;
        while (ci < ncaps && caps[ci] < n) ci++;
        if (ci == ncaps)
            return fail(ctx, STM_ERR_UNSUPPORTED, "document with more than 8192 distinct words");
        members[ci].push_back((int)d);
    }
    int max_warps = 0;
    std::vector<LengthClass> new_classes;
    std::vector<int> class_of;   // index into members[] of every new class
    for (int ci = 0; ci < ncaps; ++ci) {
        if (members[ci].empty()) continue;
        LengthClass lc;
        lc.J = caps[ci] <= 64 ? 2 : (caps[ci] <= 128 ? 4 : (caps[ci] <= 160 ? 5 : 8));
        // exact fit: tile capacity = longest document of the class (rounded to 4 rows), not the class bound
        int longest = 1;
        for (int dd : members[ci]) longest = std::max<int>(longest, (int)(doc_ptr[dd + 1] - doc_ptr[dd]));
        lc.n_cap = std::min(caps[ci], (longest + 3) / 4 * 4);
        lc.n_docs = (int)members[ci].size();
        // longest documents first: better tail balance on the dynamic queue
        std::stable_sort(members[ci].begin(), members[ci].end(), [&](int a, int b) {
            return (doc_ptr[a + 1] - doc_ptr[a]) > (doc_ptr[b + 1] - doc_ptr[b]);
        });
        lc.smem_per_warp = (int)bfgs_smem_tile(lc.n_cap, ctx->TS, ctx->KPL);
        lc.smem_small = (int)bfgs_smem_small(ctx->KPL);
        {
            // TMEM residency: a warp's lanes hold ceil(n_cap/32) word slots of CS columns each; 8 warps
            // x 256 columns or 4 warps x 512 columns (a warp can only address its own 32-lane quarter)
            const int slots = (lc.n_cap + 31) / 32, CS = (ctx->K + 1) & ~1;
            if (lc.n_cap <= 32 * lc.J) {
                if (slots * CS <= 256) { lc.tm_warps = 8; lc.tm_cols = 256; }
                else if (slots * CS <= 512) { lc.tm_warps = 4; lc.tm_cols = 512; }
            }
            const int max_w = std::min(STM_BFGS_MAX_THREADS / 32, ctx->tune_bfgs_max_warps);
            lc.tm_warps = std::min(lc.tm_warps, max_w);
            const int avail = ctx->max_smem - 128 - lc.tm_warps * lc.smem_small;   // 128: static shared (TMEM base)
            int sw = avail > 0 ? avail / (lc.smem_small + lc.smem_per_warp) : 0;
            sw = std::max(0, std::min(sw, max_w - lc.tm_warps));
            lc.warps = lc.tm_warps + sw;
        }
        lc.post_smem_per_warp = (int)post_smem_per_warp(lc.n_cap, ctx->TS, ctx->K1, ctx->KPL);
        lc.post_warps = std::min(8, ctx->max_smem / lc.post_smem_per_warp);
        if (lc.warps < 1 || lc.post_warps < 1)
            return fail(ctx, STM_ERR_UNSUPPORTED,
                        "a document's beta tile (" + std::to_string(lc.post_smem_per_warp) +
                            " bytes) does not fit in shared memory");
        lc.grid = std::min(ctx->sm_count, (lc.n_docs + lc.warps - 1) / lc.warps);
        lc.post_grid = std::min(ctx->sm_count, (lc.n_docs + lc.post_warps - 1) / lc.post_warps);
        {
            const int gw = post_group_warps(ctx->K1, ctx->KPL);
            const int per_group = (int)post_group_smem(lc.n_cap, ctx->TS, ctx->K1, ctx->KPL, gw);
            const int g = std::min(stm::post_group_max_threads(gw) / (32 * gw), ctx->max_smem / per_group);
            if (g >= 1) {
                lc.post_groups = g;
                lc.post_gw = gw;
                lc.post_smem_per_warp = per_group;
                // CTAs per SM: what the kernel's register budget was cut for, if the shared memory agrees
                const int per_sm = std::max(1, std::min(stm::post_group_min_blocks(gw), ctx->max_smem / (per_group * g)));
                lc.post_grid = std::min(ctx->sm_count * per_sm, (lc.n_docs + g - 1) / g);
                max_warps = std::max(max_warps, lc.post_grid * g);
            }
        }
        max_warps = std::max(max_warps, std::max(lc.grid * lc.warps, lc.post_grid * lc.post_warps));
        new_classes.push_back(lc);
        class_of.push_back(ci);
    }
    // everything is validated: release the previous corpus and take the new one
    free_corpus(ctx);
    ctx->kappa_warm = nullptr;
    ctx->D = D; ctx->nnz = nnz; ctx->n_max = n_max;
    // host-API chunks: equal document ranges, enough documents per chunk to fill the GPU several times over
    ctx->n_chunks = (int)std::max<int64_t>(1, std::min<int64_t>(ctx->tune_host_chunks, D / 8192));
    ctx->chunk_lo.assign(ctx->n_chunks + 1, 0);
    if (ctx->n_chunks >= 3) {
        // small first and last chunks (their copies are the ones nothing overlaps), the rest split evenly
        const int64_t edge = (int64_t)((double)D / (2.5 * ctx->n_chunks));
        const int mid = ctx->n_chunks - 2;
        for (int c = 1; c < ctx->n_chunks; ++c) ctx->chunk_lo[c] = edge + (D - 2 * edge) * (c - 1) / mid;
        ctx->chunk_lo[ctx->n_chunks] = D;
    } else {
        for (int c = 0; c <= ctx->n_chunks; ++c) ctx->chunk_lo[c] = D * c / ctx->n_chunks;
    }
    for (size_t i = 0; i < new_classes.size(); ++i) {
        LengthClass& lc = new_classes[i];
        CU(cudaMalloc(&lc.d_docs, sizeof(int) * lc.n_docs));
        CU(cudaMemcpy(lc.d_docs, members[class_of[i]].data(), sizeof(int) * lc.n_docs, cudaMemcpyHostToDevice));
        // the same documents grouped by chunk (each group keeps the longest-first order)
        std::vector<int> grouped;
        grouped.reserve(lc.n_docs);
        lc.chunk_off.assign(1, 0);
        for (int c = 0; c < ctx->n_chunks; ++c) {
            for (int dd : members[class_of[i]])
                if (dd >= ctx->chunk_lo[c] && dd < ctx->chunk_lo[c + 1]) grouped.push_back(dd);
            lc.chunk_off.push_back((int)grouped.size());
        }
        CU(cudaMalloc(&lc.d_chunk_docs, sizeof(int) * lc.n_docs));
        CU(cudaMemcpy(lc.d_chunk_docs, grouped.data(), sizeof(int) * lc.n_docs, cudaMemcpyHostToDevice));
        ctx->classes.push_back(lc);
    }
    CU(cudaMalloc(&ctx->d_doc_ptr, sizeof(long long) * (D + 1)));
    CU(cudaMemcpy(ctx->d_doc_ptr, doc_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice));
    CU(cudaMalloc(&ctx->d_word_id, sizeof(int) * std::max<int64_t>(nnz, 1)));
    CU(cudaMalloc(&ctx->d_count, sizeof(float) * std::max<int64_t>(nnz, 1)));
    if (nnz) {
        CU(cudaMemcpy(ctx->d_word_id, word_id, sizeof(int) * nnz, cudaMemcpyHostToDevice));
        CU(cudaMemcpy(ctx->d_count, count, sizeof(float) * nnz, cudaMemcpyHostToDevice));
    }
    if (aspect) {
        CU(cudaMalloc(&ctx->d_aspect, sizeof(int) * std::max<int64_t>(D, 1)));
        CU(cudaMemcpy(ctx->d_aspect, aspect, sizeof(int) * D, cudaMemcpyHostToDevice));
    }
    CU(cudaMalloc(&ctx->d_queues, sizeof(unsigned int) * 32 * (stm_ctx::MAX_CHUNKS + 1)));
    CU(cudaMalloc(&ctx->d_dbg, sizeof(unsigned long long) * 16));
    CU(cudaMemset(ctx->d_dbg, 0, sizeof(unsigned long long) * 16));
    ctx->max_warps_total = std::max(max_warps, 1);
    ctx->scratch_stride = 2LL * ctx->K1 * ctx->K1;
    CU(cudaMalloc(&ctx->d_scratch, sizeof(double) * ctx->scratch_stride * ctx->max_warps_total));
    ctx->n_rep = 16;
    CU(cudaMalloc(&ctx->d_sigma_rep, sizeof(double) * ctx->n_rep * ctx->K1 * ctx->K1));
    CU(cudaMalloc(&ctx->d_ones, sizeof(double) * std::max<int64_t>(D, 1)));
    fill_kernel<<<256, 256>>>(ctx->d_ones, D, 1.0);
    ctx->colsum_blocks = std::min(1024, (ctx->V + 63) / 64);
    CU(cudaMalloc(&ctx->d_colsum_part, sizeof(double) * (size_t)ctx->colsum_blocks * ctx->TS));
    CU(cudaDeviceSynchronize());
    return STM_OK;
}

#if STM_DBG_TIMING || STM_SLOTS_TIMING
// variant builds only (not part of include/stm_b200.h): read and reset the per-phase cycle counters
int stm_dbg_cycles(stm_ctx* ctx, unsigned long long* out8) {
    cudaMemcpy(out8, ctx->d_dbg, sizeof(unsigned long long) * 16, cudaMemcpyDeviceToHost);
    cudaMemset(ctx->d_dbg, 0, sizeof(unsigned long long) * 16);
    return 0;
}
#endif

int stm_stats_layout(const stm_ctx* ctx, int p, int64_t* offsets) {
    if (!ctx || !offsets || p < 0) return STM_ERR_INVALID;
    layout(ctx, p, offsets);
    return STM_OK;
}

int stm_prologue(stm_ctx* ctx, const double* sigma_dev, double* prior_dev, int* info_dev, void* stream) {
    if (!ctx || !sigma_dev || !prior_dev || !info_dev) return STM_ERR_INVALID;
    cudaStream_t st = (cudaStream_t)stream;
    STM_ON_DEVICE(ctx);
    const int K1 = ctx->K1;
    double* L = ctx->d_msmall;  // K1*K1
    CU(cudaMemcpyAsync(L, sigma_dev, sizeof(double) * K1 * K1, cudaMemcpyDeviceToDevice, st));
    CS(cusolverDnSetStream(ctx->cusolver, st));
    CS(cusolverDnDpotrf(ctx->cusolver, CUBLAS_FILL_MODE_LOWER, K1, L, K1, ctx->d_potrf_work, ctx->potrf_lwork, info_dev));
    prior_from_chol_kernel<<<1, 32, 0, st>>>(L, K1, info_dev, prior_dev);
    ctx->launches++;
    CU(cudaGetLastError());
    return STM_OK;
}

// device buffers of one E-step (what stm_estep is given)
struct EstepBufs {
    const float* beta_t; const double* mu; const double* prior;
    double* eta; double* theta; double* stats; double* doc_bound; int32_t* doc_info; int32_t* doc_nfev;
};

// One phase (0: kernel A, per-document BFGS; 1: kernel B, post-optimisation) over every length class, for the whole
// corpus (chunk < 0) or for the documents of one host-API chunk.
static int launch_phase(stm_ctx* ctx, const EstepBufs& B, int phase, int chunk, cudaStream_t st) {
    int64_t off[10];
    layout(ctx, 0, off);
    int ci = 0;
    for (const auto& lc : ctx->classes) {
        stm::EstepParams P;
        P.doc_ptr = ctx->d_doc_ptr; P.word_id = ctx->d_word_id; P.count = ctx->d_count; P.aspect = ctx->d_aspect;
        if (chunk < 0) { P.docs = lc.d_docs; P.n_docs = lc.n_docs; }
        else { P.docs = lc.d_chunk_docs + lc.chunk_off[chunk]; P.n_docs = lc.chunk_off[chunk + 1] - lc.chunk_off[chunk]; }
        P.queue = ctx->d_queues + 32 * (chunk + 1) + 16 * phase + ci;
        ci++;
        if (P.n_docs == 0) continue;
        P.K = ctx->K; P.V = ctx->V; P.A = ctx->A; P.TS = ctx->TS;
        P.beta_t = B.beta_t; P.mu = B.mu; P.prior = B.prior;
        P.eta = B.eta; P.theta = B.theta; P.doc_bound = B.doc_bound; P.doc_info = B.doc_info;
        P.doc_nfev = B.doc_nfev;
        P.beta_ss_t = B.stats + off[0];
        P.sigma_ss_rep = ctx->d_sigma_rep; P.n_rep = ctx->n_rep;
        P.scratch = ctx->d_scratch; P.scratch_stride = ctx->scratch_stride;
        P.n_cap = lc.n_cap; P.smem_per_warp = lc.smem_per_warp;
        P.smem_small = lc.smem_small; P.tm_warps = lc.tm_warps; P.tm_cols = lc.tm_cols;
        P.dbg_cycles = ctx->d_dbg;
        if (phase == 0) {
            const int grid = std::min(lc.grid, (P.n_docs + lc.warps - 1) / lc.warps);
            CU(launch_bfgs(ctx->KPL, P, lc.J, grid, lc.warps * 32,
                           (size_t)lc.smem_small * lc.warps + (size_t)lc.smem_per_warp * (lc.warps - lc.tm_warps), st));
        } else {
            P.smem_per_warp = lc.post_smem_per_warp;
            if (lc.post_groups > 0) {
                const size_t smem = (size_t)lc.post_smem_per_warp * lc.post_groups;
                const int block = lc.post_groups * lc.post_gw * 32;
                const int grid = std::min(lc.post_grid, (P.n_docs + lc.post_groups - 1) / lc.post_groups);
                CU(ctx->KPL == 1   ? stm_launch_post_group_kpl1(P, lc.post_gw, grid, block, smem, st)
                   : ctx->KPL == 2 ? stm_launch_post_group_kpl2(P, lc.post_gw, grid, block, smem, st)
                   : ctx->KPL == 3 ? stm_launch_post_group_kpl3(P, lc.post_gw, grid, block, smem, st)
                                   : stm_launch_post_group_kpl4(P, lc.post_gw, grid, block, smem, st));
            } else {
                const int grid = std::min(lc.post_grid, (P.n_docs + lc.post_warps - 1) / lc.post_warps);
                CU(launch_post(ctx->KPL, P, grid, lc.post_warps * 32, (size_t)lc.post_smem_per_warp * lc.post_warps, st));
            }
        }
        ctx->launches++;
    }
    return STM_OK;
}

static int estep_begin(stm_ctx* ctx, double* stats_dev, cudaStream_t st) {
    int64_t off[10];
    layout(ctx, 0, off);
    CU(cudaMemsetAsync(stats_dev + off[0], 0, sizeof(double) * (size_t)ctx->A * ctx->V * ctx->TS, st));
    CU(cudaMemsetAsync(ctx->d_sigma_rep, 0, sizeof(double) * ctx->n_rep * ctx->K1 * ctx->K1, st));
    CU(cudaMemsetAsync(ctx->d_queues, 0, sizeof(unsigned int) * 32 * (stm_ctx::MAX_CHUNKS + 1), st));
    return STM_OK;
}

static int estep_end(stm_ctx* ctx, const EstepBufs& B, cudaStream_t st) {
    int64_t off[10];
    layout(ctx, 0, off);
    estep_epilogue_kernel<<<1, 1024, 0, st>>>(ctx->d_sigma_rep, ctx->n_rep, ctx->K1, B.doc_bound, ctx->D,
                                              B.stats + off[1], B.stats + off[2], B.stats + off[3]);
    ctx->launches++;
    CU(cudaGetLastError());
    return STM_OK;
}

int stm_estep(stm_ctx* ctx, const float* beta_t_dev, const double* mu_dev, const double* prior_dev,
              double* eta_dev, double* theta_dev, double* stats_dev, double* doc_bound_dev,
              int32_t* doc_info_dev, int32_t* doc_nfev_dev, void* stream) {
    if (!ctx) return STM_ERR_INVALID;
    if (!ctx->d_doc_ptr) return fail(ctx, STM_ERR_NO_CORPUS, "stm_set_corpus has not been called");
    if (!beta_t_dev || !mu_dev || !prior_dev || !eta_dev || !theta_dev || !stats_dev || !doc_bound_dev ||
        !doc_info_dev || !doc_nfev_dev)
        return fail(ctx, STM_ERR_INVALID, "NULL device pointer passed to stm_estep");
    cudaStream_t st = (cudaStream_t)stream;
    STM_ON_DEVICE(ctx);
    const EstepBufs B{beta_t_dev, mu_dev, prior_dev, eta_dev, theta_dev, stats_dev, doc_bound_dev, doc_info_dev,
                      doc_nfev_dev};
    int rc = estep_begin(ctx, stats_dev, st);
    if (rc) return rc;
    // kernel A (BFGS) for every length class, then kernel B (post-optimisation) for every length class;
    // three events bracket the two phases (stm_estep_kernel_ms)
    CU(cudaEventRecord(ctx->ev[0], st));
    for (int phase = 0; phase < 2; ++phase) {
        rc = launch_phase(ctx, B, phase, -1, st);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev[phase + 1], st));
    }
    return estep_end(ctx, B, st);
}

int stm_moments(stm_ctx* ctx, const double* eta_dev, const double* x_dev, int p, double* stats_dev, void* stream) {
    if (!ctx || !eta_dev || !stats_dev || p < 0 || (p > 0 && !x_dev)) return STM_ERR_INVALID;
    if (!ctx->d_doc_ptr) return fail(ctx, STM_ERR_NO_CORPUS, "stm_set_corpus has not been called");
    cudaStream_t st = (cudaStream_t)stream;
    STM_ON_DEVICE(ctx);
    CB(cublasSetStream(ctx->cublas, st));
    int64_t off[10];
    layout(ctx, p, off);
    const int K1 = ctx->K1;
    const int D = (int)ctx->D;
    const double one = 1.0, zero = 0.0;
    if (D == 0) {
        CU(cudaMemsetAsync(stats_dev + off[4], 0, sizeof(double) * (off[9] - off[4]), st));
        return STM_OK;
    }
    // eta is row-major [D][K1] == column-major K1 x D (ld K1); X row-major [D][p] == column-major p x D
    CB(cublasDgemv(ctx->cublas, CUBLAS_OP_N, K1, D, &one, eta_dev, K1, ctx->d_ones, 1, &zero, stats_dev + off[4], 1));
    CB(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, K1, K1, D, &one, eta_dev, K1, eta_dev, K1, &zero,
                   stats_dev + off[8], K1));
    if (p > 0) {
        CB(cublasDgemv(ctx->cublas, CUBLAS_OP_N, p, D, &one, x_dev, p, ctx->d_ones, 1, &zero, stats_dev + off[5], 1));
        CB(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, p, p, D, &one, x_dev, p, x_dev, p, &zero,
                       stats_dev + off[6], p));
        // xte row-major [p][K1] == column-major K1 x p = eta' X'^T
        CB(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_T, K1, p, D, &one, eta_dev, K1, x_dev, p, &zero,
                       stats_dev + off[7], K1));
    }
    return STM_OK;
}

int stm_mstep(stm_ctx* ctx, const double* stats_dev, const double* x_dev, int p, int model, double sigprior,
              double* gamma_t_dev, double* mu_dev, double* sigma_dev, float* beta_t_dev, double* beta64_t_dev,
              void* stream) {
    if (!ctx || !stats_dev || !mu_dev || !sigma_dev || !beta_t_dev) return STM_ERR_INVALID;
    if (!(sigprior >= 0.0 && sigprior <= 1.0))
        return fail(ctx, STM_ERR_INVALID, "weight needs to be defined between 0 and 1");  // stm.py:721
    const int reg_mode = model;   // STM_MODEL_STM (ols) / STM_MODEL_STM_RIDGE / STM_MODEL_STM_LASSO / STM_MODEL_CTM
    if (model == STM_MODEL_STM_RIDGE || model == STM_MODEL_STM_LASSO) model = STM_MODEL_STM;
    if (model != STM_MODEL_STM && model != STM_MODEL_CTM)
        return fail(ctx, STM_ERR_INVALID, "Updating the topical prevalence parameter requires a mode");  // stm.py:709
    if (model == STM_MODEL_STM && (p < 1 || !x_dev || !gamma_t_dev))
        return fail(ctx, STM_ERR_INVALID, "STM mode needs a design matrix with p >= 1 and a gamma buffer");
    cudaStream_t st = (cudaStream_t)stream;
    STM_ON_DEVICE(ctx);
    CB(cublasSetStream(ctx->cublas, st));
    CS(cusolverDnSetStream(ctx->cusolver, st));
    const int K1 = ctx->K1, K = ctx->K, V = ctx->V, TS = ctx->TS;
    const int D = (int)ctx->D;
    int64_t off[10];
    layout(ctx, model == STM_MODEL_STM ? p : p, off);
    // small workspace carve-up
    const int64_t need = 16 + 2LL * p * p + p + 5LL * p * K1 + 2LL * TS * ctx->A;
    if (need > ctx->msmall_len) {
        cudaFree(ctx->d_msmall);
        ctx->msmall_len = need + 1024;
        CU(cudaMalloc(&ctx->d_msmall, sizeof(double) * ctx->msmall_len));
    }
    long long* d_off = reinterpret_cast<long long*>(ctx->d_msmall);  // 10 offsets in the first 16 doubles
    double* G = ctx->d_msmall + 16;
    double* lam = G + (int64_t)p * p;
    double* R = lam + p;
    double* tmp = R + (int64_t)p * K1;
    double* rowsum = tmp + (int64_t)p * K1;
    double* lasso_ws = rowsum + 2LL * TS * ctx->A;     // lasso: 3 p K1
    long long hoff[10];
    for (int i = 0; i < 10; ++i) hoff[i] = off[i];
    CU(cudaMemcpyAsync(d_off, hoff, sizeof(hoff), cudaMemcpyHostToDevice, st));

    // ---- update_mu ---------------------------------------------------------------------------
    if (model == STM_MODEL_STM) {
        center_moments_kernel<<<1, 256, 0, st>>>(stats_dev, d_off, p, K1, G, R);
        if (reg_mode == STM_MODEL_STM_LASSO) {
            lasso_gamma_kernel<<<(K1 + 63) / 64, 64, 0, st>>>(G, R, stats_dev, d_off, p, K1, 1.0, lasso_ws, gamma_t_dev);
            ctx->launches += 2;
        } else {
        if (ctx->syevd_p != p) {
            int lwork = 0;
            CS(cusolverDnDsyevd_bufferSize(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, p, G, p,
                                           lam, &lwork));
            cudaFree(ctx->d_syevd_work);
            ctx->syevd_lwork = std::max(lwork, 1);
            CU(cudaMalloc(&ctx->d_syevd_work, sizeof(double) * ctx->syevd_lwork));
            ctx->syevd_p = p;
        }
        CS(cusolverDnDsyevd(ctx->cusolver, CUSOLVER_EIG_MODE_VECTOR, CUBLAS_FILL_MODE_LOWER, p, G, p, lam,
                            ctx->d_syevd_work, ctx->syevd_lwork, ctx->d_info + 1));
        solve_gamma_kernel<<<1, 256, 0, st>>>(G, lam, R, p, K1, reg_mode == STM_MODEL_STM_RIDGE ? 0.1 : 0.0, tmp,
                                              gamma_t_dev);
        ctx->launches += 2;
        }
        // mu (column-major K1 x D) = gamma_t (column-major K1 x p) * X (column-major p x D)
        if (D > 0) {
            const double one = 1.0, zero = 0.0;
            CB(cublasDgemm(ctx->cublas, CUBLAS_OP_N, CUBLAS_OP_N, K1, D, p, &one, gamma_t_dev, K1, x_dev, p, &zero,
                           mu_dev, K1));
        }
    } else {
        mu_ctm_kernel<<<256, 256, 0, st>>>(stats_dev, d_off, K1, ctx->D, mu_dev);
        ctx->launches++;
    }
    // ---- update_sigma ------------------------------------------------------------------------
    sigma_update_kernel<<<1, 256, 0, st>>>(stats_dev, d_off, p, K1, model, gamma_t_dev, sigprior, tmp, sigma_dev);
    ctx->launches++;
    // ---- update_beta -------------------------------------------------------------------------
    const double* ss = stats_dev + off[0];
    if (ctx->A == 1) {
        const int rows_per_block = (V + ctx->colsum_blocks - 1) / ctx->colsum_blocks;
        const int nblk = (V + rows_per_block - 1) / rows_per_block;
        colsum_partial_kernel<<<nblk, ((TS + 31) / 32) * 32, 0, st>>>(ss, V, TS, rows_per_block, ctx->d_colsum_part);
        colsum_final_kernel<<<1, ((TS + 31) / 32) * 32, 0, st>>>(ctx->d_colsum_part, nblk, TS, rowsum);
        beta_normalise_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(ss, rowsum, V, K, TS, beta_t_dev, beta64_t_dev);
        ctx->launches += 3;
    } else {
        beta_normalise_aspect_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(ss, (long long)ctx->A * V, K, TS, beta_t_dev,
                                                                        beta64_t_dev);
        ctx->launches++;
    }
    CU(cudaGetLastError());
    return STM_OK;
}

int stm_beta_to_wordmajor(stm_ctx* ctx, const double* beta_kv_dev, float* beta_t_dev, void* stream) {
    if (!ctx || !beta_kv_dev || !beta_t_dev) return STM_ERR_INVALID;
    STM_ON_DEVICE(ctx);
    beta_to_wordmajor_kernel<<<ctx->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(beta_kv_dev, beta_t_dev, ctx->A,
                                                                                  ctx->K, ctx->V, ctx->TS);
    ctx->launches++;
    CU(cudaGetLastError());
    return STM_OK;
}
int stm_wordmajor_to_kv(stm_ctx* ctx, const double* src_t_dev, double* dst_kv_dev, void* stream) {
    if (!ctx || !src_t_dev || !dst_kv_dev) return STM_ERR_INVALID;
    STM_ON_DEVICE(ctx);
    wordmajor_to_kv_kernel<<<ctx->sm_count * 4, 256, 0, (cudaStream_t)stream>>>(src_t_dev, dst_kv_dev, ctx->A, ctx->K,
                                                                                ctx->V, ctx->TS);
    ctx->launches++;
    CU(cudaGetLastError());
    return STM_OK;
}

int stm_estep_host(stm_ctx* ctx, const double* beta, const double* mu, const double* siginv, double sigmaentropy,
                   double* eta, double* theta, double* beta_ss, double* sigma_ss, double* bound, double* doc_bound,
                   int32_t* doc_status, int32_t* doc_nit, int32_t* doc_repair) {
    if (!ctx) return STM_ERR_INVALID;
    if (!ctx->d_doc_ptr) return fail(ctx, STM_ERR_NO_CORPUS, "stm_set_corpus has not been called");
    if (!beta || !mu || !siginv || !eta || !theta || !beta_ss || !sigma_ss || !bound)
        return fail(ctx, STM_ERR_INVALID, "NULL host pointer passed to stm_estep_host");
    STM_ON_DEVICE(ctx);
    const int K = ctx->K, K1 = ctx->K1, V = ctx->V, A = ctx->A, TS = ctx->TS;
    const int64_t D = ctx->D;
    // the reference's siginv is diagonal by construction (stm.py:501: element-wise product of a lower- and
    // an upper-triangular inverse); anything else is not this path.  When LAPACK's LU pivots inside
    // np.linalg.inv(chol) the "zero" triangle holds rounding noise (~1e-17), so off-diagonals are accepted
    // up to 1e-9 of the geometric mean of their diagonal entries (their effect is below fp64 resolution).
    std::vector<double> prior(K1 + 1);
    for (int i = 0; i < K1; ++i)
        for (int j = 0; j < K1; ++j) {
            const double s = siginv[(size_t)i * K1 + j];
            if (i == j) prior[i] = s;
            else if (!(std::fabs(s) <= 1e-9 * std::sqrt(std::fabs(siginv[(size_t)i * K1 + i] * siginv[(size_t)j * K1 + j]))))
                return fail(ctx, STM_ERR_UNSUPPORTED, "siginv must be diagonal (as produced by stm.py:501)");
        }
    prior[K1] = sigmaentropy;
    int64_t off[10];
    layout(ctx, 0, off);
    if (!ctx->host_bufs) {
        const size_t akv = (size_t)A * K * V;
        CU(cudaMalloc(&ctx->h_beta_kv, sizeof(double) * akv));
        CU(cudaMalloc(&ctx->h_bss_kv, sizeof(double) * akv));
        CU(cudaMalloc(&ctx->h_beta_t, sizeof(float) * (size_t)A * V * TS));
        CU(cudaMalloc(&ctx->h_mu, sizeof(double) * std::max<int64_t>(D * K1, 1)));
        CU(cudaMalloc(&ctx->h_eta, sizeof(double) * std::max<int64_t>(D * K1, 1)));
        CU(cudaMalloc(&ctx->h_theta, sizeof(double) * std::max<int64_t>(D * K, 1)));
        CU(cudaMalloc(&ctx->h_stats, sizeof(double) * off[9]));
        CU(cudaMalloc(&ctx->h_prior, sizeof(double) * (K1 + 1)));
        CU(cudaMalloc(&ctx->h_doc_bound, sizeof(double) * std::max<int64_t>(D, 1)));
        CU(cudaMalloc(&ctx->h_doc_info, sizeof(int32_t) * std::max<int64_t>(3 * D, 1)));
        CU(cudaMalloc(&ctx->h_doc_nfev, sizeof(int32_t) * std::max<int64_t>(D, 1)));
        ctx->host_bufs = true;
    }
    // Three streams: inputs go up on `in`, the kernels run on `st`, results come home on `out`.  The documents are
    // split into index chunks (stm_set_corpus); chunk c's mu / eta are copied while chunk c-1 computes, its eta goes
    // home as soon as its kernel A is done and its theta / per-document outputs as soon as its kernel B is done.
    cudaStream_t st = ctx->host_stream, in = ctx->in_stream, out = ctx->copy_stream;
    const int NC = ctx->n_chunks;
    CU(cudaMemcpyAsync(ctx->h_beta_kv, beta, sizeof(double) * (size_t)A * K * V, cudaMemcpyHostToDevice, in));
    CU(cudaMemcpyAsync(ctx->h_prior, prior.data(), sizeof(double) * (K1 + 1), cudaMemcpyHostToDevice, in));
    CU(cudaEventRecord(ctx->ev_misc[0], in));
    for (int c = 0; c < NC; ++c) {
        const int64_t lo = ctx->chunk_lo[c], n = ctx->chunk_lo[c + 1] - lo;
        if (n > 0) {
            CU(cudaMemcpyAsync(ctx->h_mu + lo * K1, mu + lo * K1, sizeof(double) * n * K1, cudaMemcpyHostToDevice, in));
            CU(cudaMemcpyAsync(ctx->h_eta + lo * K1, eta + lo * K1, sizeof(double) * n * K1, cudaMemcpyHostToDevice, in));
        }
        CU(cudaEventRecord(ctx->ev_in[c], in));
    }
    CU(cudaStreamWaitEvent(st, ctx->ev_misc[0], 0));
    int rc = stm_beta_to_wordmajor(ctx, ctx->h_beta_kv, ctx->h_beta_t, st);
    if (rc) return rc;
    const EstepBufs B{ctx->h_beta_t, ctx->h_mu, ctx->h_prior, ctx->h_eta, ctx->h_theta, ctx->h_stats, ctx->h_doc_bound,
                      ctx->h_doc_info, ctx->h_doc_nfev};
    rc = estep_begin(ctx, ctx->h_stats, st);
    if (rc) return rc;
    for (int c = 0; c < NC; ++c) {
        const int64_t lo = ctx->chunk_lo[c], n = ctx->chunk_lo[c + 1] - lo;
        CU(cudaStreamWaitEvent(st, ctx->ev_in[c], 0));
        rc = launch_phase(ctx, B, 0, c, st);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_a[c], st));
        rc = launch_phase(ctx, B, 1, c, st);
        if (rc) return rc;
        CU(cudaEventRecord(ctx->ev_b[c], st));
        if (n > 0) {
            CU(cudaStreamWaitEvent(out, ctx->ev_a[c], 0));
            CU(cudaMemcpyAsync(eta + lo * K1, ctx->h_eta + lo * K1, sizeof(double) * n * K1, cudaMemcpyDeviceToHost, out));
            CU(cudaStreamWaitEvent(out, ctx->ev_b[c], 0));
            CU(cudaMemcpyAsync(theta + lo * K, ctx->h_theta + lo * K, sizeof(double) * n * K, cudaMemcpyDeviceToHost, out));
            if (doc_bound)
                CU(cudaMemcpyAsync(doc_bound + lo, ctx->h_doc_bound + lo, sizeof(double) * n, cudaMemcpyDeviceToHost, out));
        }
    }
    rc = estep_end(ctx, B, st);
    if (rc) return rc;
    rc = stm_wordmajor_to_kv(ctx, ctx->h_stats + off[0], ctx->h_bss_kv, st);
    if (rc) return rc;
    // per-document diagnostics: unpacked on the compute stream into the two spare thirds of the info buffer
    int32_t* u = ctx->h_doc_info;
    if (doc_status) unpack_info_kernel<<<256, 256, 0, st>>>(ctx->h_doc_info, D, u + D, nullptr, nullptr);
    if (doc_nit) unpack_info_kernel<<<256, 256, 0, st>>>(ctx->h_doc_info, D, nullptr, u + 2 * D, nullptr);
    CU(cudaEventRecord(ctx->ev_misc[1], st));
    CU(cudaStreamWaitEvent(out, ctx->ev_misc[1], 0));
    CU(cudaMemcpyAsync(beta_ss, ctx->h_bss_kv, sizeof(double) * (size_t)A * K * V, cudaMemcpyDeviceToHost, out));
    CU(cudaMemcpyAsync(sigma_ss, ctx->h_stats + off[1], sizeof(double) * K1 * K1, cudaMemcpyDeviceToHost, out));
    CU(cudaMemcpyAsync(bound, ctx->h_stats + off[2], sizeof(double), cudaMemcpyDeviceToHost, out));
    if (doc_status) CU(cudaMemcpyAsync(doc_status, u + D, sizeof(int32_t) * D, cudaMemcpyDeviceToHost, out));
    if (doc_nit) CU(cudaMemcpyAsync(doc_nit, u + 2 * D, sizeof(int32_t) * D, cudaMemcpyDeviceToHost, out));
    if (doc_repair) {
        // reuses the status third: after the status copy (same stream order on `out`)
        CU(cudaEventRecord(ctx->ev_misc[0], out));
        CU(cudaStreamWaitEvent(st, ctx->ev_misc[0], 0));
        unpack_info_kernel<<<256, 256, 0, st>>>(ctx->h_doc_info, D, nullptr, nullptr, u + D);
        CU(cudaEventRecord(ctx->ev_misc[1], st));
        CU(cudaStreamWaitEvent(out, ctx->ev_misc[1], 0));
        CU(cudaMemcpyAsync(doc_repair, u + D, sizeof(int32_t) * D, cudaMemcpyDeviceToHost, out));
    }
    CU(cudaStreamSynchronize(out));
    CU(cudaStreamSynchronize(st));
    return STM_OK;
}

int stm_heldout(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr_dev, const int32_t* word_id_dev,
                const float* count_dev, const double* theta_dev, const float* beta_t_dev, double* doc_ll_dev,
                double* mean_dev, void* stream) {
    if (!ctx) return STM_ERR_INVALID;
    if (D < 1 || !doc_ptr_dev || !word_id_dev || !count_dev || !theta_dev || !beta_t_dev || !doc_ll_dev || !mean_dev)
        return fail(ctx, STM_ERR_INVALID, "stm_heldout: NULL pointer or no documents");
    if (ctx->A != 1) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_heldout: eval_heldout takes one K x V beta (A = 1)");
    STM_ON_DEVICE(ctx);
    return heldout_launch<float>(ctx, D, doc_ptr_dev, word_id_dev, count_dev, theta_dev, beta_t_dev, doc_ll_dev,
                                 mean_dev, (cudaStream_t)stream);
}

int stm_heldout64(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr_dev, const int32_t* word_id_dev,
                  const float* count_dev, const double* theta_dev, const double* beta64_t_dev, double* doc_ll_dev,
                  double* mean_dev, void* stream) {
    if (!ctx) return STM_ERR_INVALID;
    if (D < 1 || !doc_ptr_dev || !word_id_dev || !count_dev || !theta_dev || !beta64_t_dev || !doc_ll_dev || !mean_dev)
        return fail(ctx, STM_ERR_INVALID, "stm_heldout64: NULL pointer or no documents");
    if (ctx->A != 1) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_heldout64: eval_heldout takes one K x V beta (A = 1)");
    STM_ON_DEVICE(ctx);
    return heldout_launch<double>(ctx, D, doc_ptr_dev, word_id_dev, count_dev, theta_dev, beta64_t_dev, doc_ll_dev,
                                  mean_dev, (cudaStream_t)stream);
}

int stm_heldout_host(stm_ctx* ctx, int64_t D, const int64_t* doc_ptr, const int32_t* word_id, const float* count,
                     const double* theta, const double* beta_kv, double* doc_ll, double* mean) {
    if (!ctx) return STM_ERR_INVALID;
    if (D < 1 || !doc_ptr || !theta || !beta_kv || !mean)
        return fail(ctx, STM_ERR_INVALID, "stm_heldout_host: NULL pointer or no documents");
    if (ctx->A != 1) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_heldout_host: eval_heldout takes one K x V beta (A = 1)");
    const int64_t nnz = doc_ptr[D];
    if (doc_ptr[0] != 0 || nnz < 0 || (nnz > 0 && (!word_id || !count)))
        return fail(ctx, STM_ERR_INVALID, "stm_heldout_host: bad CSR arrays");
    for (int64_t d = 0; d < D; ++d)
        if (doc_ptr[d + 1] < doc_ptr[d]) return fail(ctx, STM_ERR_INVALID, "doc_ptr must be non-decreasing");
    for (int64_t i = 0; i < nnz; ++i)
        if (word_id[i] < 0 || word_id[i] >= ctx->V) return fail(ctx, STM_ERR_INVALID, "word id out of range [0, V)");
    STM_ON_DEVICE(ctx);
    const int K = ctx->K, V = ctx->V, TS = ctx->TS;
    long long* d_ptr = nullptr; int* d_ids = nullptr; float* d_cnt = nullptr;
    double *d_theta = nullptr, *d_bkv = nullptr, *d_bt = nullptr, *d_ll = nullptr;
    int rc = STM_OK;
    auto cleanup = [&]() {
        cudaFree(d_ptr); cudaFree(d_ids); cudaFree(d_cnt); cudaFree(d_theta); cudaFree(d_bkv); cudaFree(d_bt); cudaFree(d_ll);
    };
#define HCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
        }                                                                                            \
    } while (0)
    HCU(cudaMalloc(&d_ptr, sizeof(long long) * (D + 1)));
    HCU(cudaMalloc(&d_ids, sizeof(int) * std::max<int64_t>(nnz, 1)));
    HCU(cudaMalloc(&d_cnt, sizeof(float) * std::max<int64_t>(nnz, 1)));
    HCU(cudaMalloc(&d_theta, sizeof(double) * D * K));
    HCU(cudaMalloc(&d_bkv, sizeof(double) * (size_t)K * V));
    HCU(cudaMalloc(&d_bt, sizeof(double) * (size_t)V * TS));
    HCU(cudaMalloc(&d_ll, sizeof(double) * (D + 1)));
    HCU(cudaMemcpy(d_ptr, doc_ptr, sizeof(long long) * (D + 1), cudaMemcpyHostToDevice));
    if (nnz) {
        HCU(cudaMemcpy(d_ids, word_id, sizeof(int) * nnz, cudaMemcpyHostToDevice));
        HCU(cudaMemcpy(d_cnt, count, sizeof(float) * nnz, cudaMemcpyHostToDevice));
    }
    HCU(cudaMemcpy(d_theta, theta, sizeof(double) * D * K, cudaMemcpyHostToDevice));
    HCU(cudaMemcpy(d_bkv, beta_kv, sizeof(double) * (size_t)K * V, cudaMemcpyHostToDevice));
    beta_kv_to_wordmajor_f64_kernel<<<ctx->sm_count * 4, 256>>>(d_bkv, d_bt, K, V, TS);
    ctx->launches++;
    // the host entry keeps beta in fp64 (reference arithmetic); the device entry uses the fit's fp32 beta
    rc = heldout_launch<double>(ctx, D, reinterpret_cast<const int64_t*>(d_ptr), d_ids, d_cnt, d_theta, d_bt, d_ll,
                                d_ll + D, nullptr);
    if (rc == STM_OK) {
        HCU(cudaMemcpy(mean, d_ll + D, sizeof(double), cudaMemcpyDeviceToHost));
        if (doc_ll) HCU(cudaMemcpy(doc_ll, d_ll, sizeof(double) * D, cudaMemcpyDeviceToHost));
    }
#undef HCU
    cleanup();
    return rc;
}

}  // extern "C"

#include "spectral.cuh"
#include "corpus_gen.cuh"
#include "mnreg.cuh"
