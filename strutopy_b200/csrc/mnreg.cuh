// mnreg.cuh — content-covariate update of beta on the device.
// Replaces STM.mnreg, /root/reference/src/modules/stm.py:749-853 (update_beta with lda_beta=False,
// stm.py:746-747).  Included by stm_b200.cu.
//
// Per word v the reference fits sklearn PoissonRegressor(fit_intercept=False, alpha=250) of the (A K) counts
// beta_ss[a][k][v] on one-hot topic / aspect / interaction covariates (no offset), i.e. it minimises
//     f(w) = (1/n) sum_r (exp(eta_r) - y_r eta_r) + alpha/2 |w|^2,   eta_r = t_k(r) + a_a(r) + i_r,  n = A K
// — strictly convex, so the minimiser is unique; sklearn's lbfgs stops at a 1e-5 gradient, this kernel runs
// damped Newton to ~1e-12 (oracle/mnreg_numpy.py holds both; they agree to ~1e-8).  One warp per word.
// The Newton system H d = -g, H = X'CX + alpha I with C = diag(mu/n), is solved in O(n) through
//     H^-1 = (1/alpha) (I - X' (alpha C^-1 + X X')^-1 X),    X X' = I + P_topic + P_aspect
// (diagonal + rank K+A, eliminated topic block first, then a dense A x A Schur complement).
// Then kappa = coefficients (stm.py:841), beta = softmax_v(m_v + eta_r(v)) split by aspect (stm.py:847-853).
#pragma once

namespace stm_mnreg {

constexpr int AMAX = 8;   // content levels supported by the Schur solve

__device__ __forceinline__ double wsum(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ double wmax(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v = fmax(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// shared memory per warp (doubles): y n | wi n | mu n | e n | b n | z n | t K | gt K | q K | tau K  -> 6 n + 4 K
__global__ void kappa_newton_kernel(const double* __restrict__ beta_ss_t, int A, int K, int V, int TS, double alpha,
                                    int word_column, double* __restrict__ lin /* [n][V]: eta_r(v) */,
                                    double* __restrict__ kappa /* [p][V] or null */, int warm, int* __restrict__ flag) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const int n = A * K;
    double* y = sm + (size_t)wib * (6 * n + 4 * K);
    double* wi = y + n;
    double* mu = wi + n;
    double* e = mu + n;
    double* b = e + n;
    double* z = b + n;
    double* t = z + n;
    double* gt = t + K;
    double* q = gt + K;
    double* tau = q + K;
    const double inv_n = 1.0 / n;

    for (int v = blockIdx.x * wpb + wib; v < V; v += gridDim.x * wpb) {
        const int col = word_column >= 0 ? word_column : v;
        double ymax = 0.0;
        for (int r = lane; r < n; r += 32) {
            const int a = r / K, k = r - a * K;
            const double yy = beta_ss_t[((size_t)a * V + col) * TS + k];
            y[r] = yy; wi[r] = 0.0;
            ymax = fmax(ymax, fabs(yy));
        }
        for (int k = lane; k < K; k += 32) t[k] = 0.0;
        double av[AMAX], ga[AMAX];
#pragma unroll
        for (int a = 0; a < AMAX; ++a) av[a] = 0.0;
        if (warm) {
            // start from the previous M-step's coefficients (the statistics move little between EM iterations): the
            // strictly convex problem has one minimiser, reached in 1-3 Newton steps instead of 5-9 from zero
            for (int r = lane; r < n; r += 32) wi[r] = kappa[(size_t)(K + A + 1 + r) * V + v];
            for (int k = lane; k < K; k += 32) t[k] = kappa[(size_t)k * V + v];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) if (a < A) av[a] = kappa[(size_t)(K + 1 + a) * V + v];
        }
        ymax = wmax(ymax);
        const double gtol = 1e-12 * (1.0 + ymax * inv_n);
        __syncwarp();

        // objective at (t, av, wi) + s * (dt, da, di); direction held in q (topics), ga (aspects), z (rows)
        auto objective = [&](double s, bool use_dir) {
            double acc = 0.0;
            for (int r = lane; r < n; r += 32) {
                const int a = r / K, k = r - a * K;
                double aa = 0.0, da = 0.0;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) { aa = av[c]; da = ga[c]; }
                const double ii = wi[r] + (use_dir ? s * z[r] : 0.0);
                const double tt = t[k] + (use_dir ? s * q[k] : 0.0);
                aa += use_dir ? s * da : 0.0;
                const double eta = tt + aa + ii;
                acc += (exp(eta) - y[r] * eta) * inv_n + 0.5 * alpha * ii * ii;
            }
            for (int k = lane; k < K; k += 32) {
                const double tt = t[k] + (use_dir ? s * q[k] : 0.0);
                acc += 0.5 * alpha * tt * tt;
            }
            acc = wsum(acc);
#pragma unroll
            for (int c = 0; c < AMAX; ++c) {
                const double aa = av[c] + (use_dir ? s * ga[c] : 0.0);
                acc += 0.5 * alpha * aa * aa;   // zero for c >= A
            }
            return acc;
        };

        bool converged = false;
        double gmax_last = INFINITY;
        for (int it = 0; it < 100 && !converged; ++it) {
            // mu, gradient
            double gmax = 0.0;
            double gal[AMAX];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) gal[a] = 0.0;
            for (int k = lane; k < K; k += 32) gt[k] = 0.0;
            __syncwarp();
            for (int r = lane; r < n; r += 32) {
                const int a = r / K, k = r - a * K;
                double aa = 0.0;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) aa = av[c];
                const double m_ = exp(t[k] + aa + wi[r]);
                mu[r] = m_;
                const double res = (m_ - y[r]) * inv_n;
                b[r] = res;                       // residual, then X g
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) gal[c] += res;
            }
            __syncwarp();
            for (int k = lane; k < K; k += 32) {
                double s = 0.0;
                for (int a = 0; a < A; ++a) s += b[a * K + k];
                gt[k] = s + alpha * t[k];
                gmax = fmax(gmax, fabs(gt[k]));
            }
#pragma unroll
            for (int a = 0; a < AMAX; ++a) {
                ga[a] = (a < A) ? wsum(gal[a]) + alpha * av[a] : 0.0;
                gmax = fmax(gmax, fabs(ga[a]));
            }
            __syncwarp();
            // g_i = res + alpha wi (kept in z for now); b <- X g = g_t[k] + g_a[a] + g_i
            for (int r = lane; r < n; r += 32) {
                const int a = r / K, k = r - a * K;
                double gaa = 0.0;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) gaa = ga[c];
                const double gi = b[r] + alpha * wi[r];
                gmax = fmax(gmax, fabs(gi));
                z[r] = gi;
                e[r] = 1.0 / (alpha * n / mu[r] + 1.0);      // 1 / E_r,  E = alpha C^-1 + I
                b[r] = gt[k] + gaa + gi;
            }
            gmax = wmax(gmax);
            gmax_last = gmax;
            if (!(gmax > gtol)) { converged = true; break; }
            __syncwarp();
            // (E + U U') zz = b :  tau_k, sigma_a, rhs
            double sig[AMAX], rA[AMAX], S[AMAX][AMAX];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) {
                sig[a] = 0.0; rA[a] = 0.0;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) S[a][c] = 0.0;
            }
            for (int k = lane; k < K; k += 32) {
                double tk = 0.0, rT = 0.0;
                for (int a = 0; a < A; ++a) { tk += e[a * K + k]; rT += e[a * K + k] * b[a * K + k]; }
                tau[k] = tk;
                q[k] = rT;                       // beta_T for now
                const double inv = 1.0 / (1.0 + tk);
#pragma unroll
                for (int a = 0; a < AMAX; ++a) {
                    if (a < A) {
                        const double ca = e[a * K + k];
                        sig[a] += ca;
                        rA[a] += ca * b[a * K + k] - ca * rT * inv;
#pragma unroll
                        for (int c = 0; c < AMAX; ++c)
                            if (c < A) S[a][c] -= ca * e[c * K + k] * inv;
                    }
                }
            }
#pragma unroll
            for (int a = 0; a < AMAX; ++a) {
                sig[a] = wsum(sig[a]); rA[a] = wsum(rA[a]);
#pragma unroll
                for (int c = 0; c < AMAX; ++c) S[a][c] = wsum(S[a][c]);
                S[a][a] += 1.0 + sig[a];          // rows a >= A become the identity
            }
            // dense A x A solve (Gaussian elimination, SPD): qA
            double qA[AMAX];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) {
                const double piv = 1.0 / S[a][a];
#pragma unroll
                for (int r2 = a + 1; r2 < AMAX; ++r2) {
                    const double f_ = S[r2][a] * piv;
#pragma unroll
                    for (int c = a; c < AMAX; ++c) S[r2][c] -= f_ * S[a][c];
                    rA[r2] -= f_ * rA[a];
                }
            }
#pragma unroll
            for (int a = AMAX - 1; a >= 0; --a) {
                double s = rA[a];
#pragma unroll
                for (int c = a + 1; c < AMAX; ++c) s -= S[a][c] * qA[c];
                qA[a] = s / S[a][a];
            }
            __syncwarp();
            for (int k = lane; k < K; k += 32) {
                double s = q[k];
                for (int a = 0; a < A; ++a) {
                    double qa = 0.0;
#pragma unroll
                    for (int c = 0; c < AMAX; ++c) if (c == a) qa = qA[c];
                    s -= e[a * K + k] * qa;
                }
                tau[k] = s / (1.0 + tau[k]);      // q_T[k]
            }
            __syncwarp();
            // zz_r = e_r (b_r - q_T[k] - q_A[a]);  direction d = -(g - X' zz) / alpha
            double dA[AMAX];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) dA[a] = 0.0;
            for (int r = lane; r < n; r += 32) {
                const int a = r / K, k = r - a * K;
                double qa = 0.0;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) qa = qA[c];
                const double zz = e[r] * (b[r] - tau[k] - qa);
                b[r] = zz;
#pragma unroll
                for (int c = 0; c < AMAX; ++c) if (c == a) dA[c] += zz;
                z[r] = -(z[r] - zz) / alpha;      // d_i
            }
            __syncwarp();
            for (int k = lane; k < K; k += 32) {
                double s = 0.0;
                for (int a = 0; a < A; ++a) s += b[a * K + k];
                q[k] = -(gt[k] - s) / alpha;      // d_t
            }
            double slope = 0.0;
#pragma unroll
            for (int a = 0; a < AMAX; ++a) {
                const double g_a = ga[a];
                ga[a] = (a < A) ? -(g_a - wsum(dA[a])) / alpha : 0.0;   // d_a (ga now holds the direction)
                slope += g_a * ga[a];
            }
            __syncwarp();
            double sl = 0.0;
            for (int r = lane; r < n; r += 32) sl += ((mu[r] - y[r]) * inv_n + alpha * wi[r]) * z[r];
            for (int k = lane; k < K; k += 32) sl += gt[k] * q[k];
            slope += wsum(sl);
            // backtracking (Armijo) on f
            const double f0 = objective(0.0, false);
            double s = 1.0;
            // A Newton decrement below the resolution of f (n-term sums: ~1e-12 relative) cannot be tested through f:
            // an Armijo test on rounding noise either rejects every step or accepts a tiny one and stalls (seen at
            // 500k documents: 70 of 20 000 words took 100 noise-sized steps).  Newton's iteration converges
            // quadratically there, so the full step is taken.
            bool ok = (-slope <= 1e-10 * (1.0 + fabs(f0)));
            for (int bt = 0; bt < 50 && !ok; ++bt) {
                const double f1 = objective(s, true);
                // Armijo, with the decrease allowed to drown in the rounding of f near the minimiser (so that
                // full Newton steps keep being taken until the GRADIENT test above stops the iteration)
                if (f1 <= f0 + 1e-4 * s * slope + 8.9e-16 * fabs(f0)) { ok = true; break; }
                s *= 0.5;
            }
            if (!ok) { converged = true; break; }   // no descent at fp64 resolution: at the minimiser
            for (int r = lane; r < n; r += 32) wi[r] += s * z[r];
            for (int k = lane; k < K; k += 32) t[k] += s * q[k];
#pragma unroll
            for (int a = 0; a < AMAX; ++a) av[a] += s * ga[a];
            __syncwarp();
        }
        // not converged to the 1e-12 target after 100 iterations is an error only if the gradient is not even 1000x below
        // the 1e-5 at which the reference's solver (sklearn lbfgs, tol=1e-5) stops
        if (!converged && !(gmax_last <= 1e-8 * (1.0 + ymax * inv_n)) && lane == 0) { atomicAdd(flag, 1); atomicExch(flag + 1, v); }
        // outputs: eta_r(v) for the softmax pass; kappa rows (stm.py:769-793 column layout: topics 0..K-1,
        // [K empty], aspects K+1..K+A, interactions K+A+1..K+A+n)
        for (int r = lane; r < n; r += 32) {
            const int a = r / K, k = r - a * K;
            double aa = 0.0;
#pragma unroll
            for (int c = 0; c < AMAX; ++c) if (c == a) aa = av[c];
            lin[(size_t)r * V + v] = t[k] + aa + wi[r];
            if (kappa) kappa[(size_t)(K + A + 1 + r) * V + v] = wi[r];
        }
        if (kappa) {
            for (int k = lane; k < K; k += 32) kappa[(size_t)k * V + v] = t[k];
            if (lane == 0) {
                kappa[(size_t)K * V + v] = 0.0;
#pragma unroll
                for (int a = 0; a < AMAX; ++a) if (a < A) kappa[(size_t)(K + 1 + a) * V + v] = av[a];
            }
        }
        __syncwarp();
    }
}

// beta[r][:] = exp(m + eta_r) / sum_v exp(m_v + eta_r(v))  (stm.py:847-850): one CTA per row, fixed-order sum
__global__ void kappa_softmax_kernel(const double* __restrict__ lin, const double* __restrict__ logm, int A, int K, int V,
                                     int TS, float* __restrict__ beta_t, double* __restrict__ beta64_t) {
    __shared__ double s[256];
    const int r = blockIdx.x, a = r / K, k = r - a * K;
    double acc = 0.0;
    for (int v = threadIdx.x; v < V; v += 256) acc += exp(logm[v] + lin[(size_t)r * V + v]);
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    const double tot = s[0];
    for (int v = threadIdx.x; v < V; v += 256) {
        const double bv = exp(logm[v] + lin[(size_t)r * V + v]) / tot;
        beta_t[((size_t)a * V + v) * TS + k] = (float)bv;
        if (beta64_t) beta64_t[((size_t)a * V + v) * TS + k] = bv;
    }
}

}  // namespace stm_mnreg

extern "C" {

int stm_update_kappa(stm_ctx* ctx, const double* stats_dev, const double* logm_dev, double alpha, int word_column,
                     float* beta_t_dev, double* beta64_t_dev, double* kappa_dev, void* stream) {
    using namespace stm_mnreg;
    if (!ctx) return STM_ERR_INVALID;
    if (!stats_dev || !logm_dev || !beta_t_dev) return fail(ctx, STM_ERR_INVALID, "stm_update_kappa: NULL pointer");
    if (!(alpha > 0.0)) return fail(ctx, STM_ERR_INVALID, "stm_update_kappa: alpha must be > 0");
    const int A = ctx->A, K = ctx->K, V = ctx->V, TS = ctx->TS, n = A * K;
    if (A < 2) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_update_kappa: the content model needs A >= 2 aspects");
    if (A > AMAX) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_update_kappa: more than 8 content levels");
    if (word_column >= V) return fail(ctx, STM_ERR_INVALID, "stm_update_kappa: word_column out of range");
    STM_ON_DEVICE(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    double* lin = nullptr; int* d_flag = nullptr;
    auto cleanup = [&]() { cudaFree(d_flag); };
#define KCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
        }                                                                                            \
    } while (0)
    if (ctx->kappa_lin_len < (int64_t)n * V) {      // [A K][V] linear predictors: kept with the context
        cudaFree(ctx->d_kappa_lin);
        ctx->d_kappa_lin = nullptr; ctx->kappa_lin_len = 0;
        KCU(cudaMalloc(&ctx->d_kappa_lin, sizeof(double) * (size_t)n * V));
        ctx->kappa_lin_len = (int64_t)n * V;
    }
    lin = ctx->d_kappa_lin;
    KCU(cudaMalloc(&d_flag, sizeof(int) * 2));
    KCU(cudaMemsetAsync(d_flag, 0, sizeof(int) * 2, st));
    const size_t per_warp = sizeof(double) * (6 * (size_t)n + 4 * (size_t)K);
    int wpb = (int)std::min<size_t>(8, ((size_t)ctx->max_smem - 1024) / per_warp);
    if (wpb < 1) { cleanup(); return fail(ctx, STM_ERR_UNSUPPORTED, "stm_update_kappa: A*K too large for shared memory"); }
    const size_t smem = per_warp * wpb;
    KCU(cudaFuncSetAttribute(kappa_newton_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
    int64_t off[10];
    layout(ctx, 0, off);
    const int grid = std::min((V + wpb - 1) / wpb, ctx->sm_count * 8);
    const int warm = (kappa_dev != nullptr && ctx->kappa_warm == kappa_dev && ctx->kappa_warm_column == word_column) ? 1 : 0;
    ctx->kappa_warm = nullptr;
    kappa_newton_kernel<<<grid, wpb * 32, smem, st>>>(stats_dev + off[0], A, K, V, TS, alpha, word_column, lin, kappa_dev,
                                                      warm, d_flag);
    kappa_softmax_kernel<<<n, 256, 0, st>>>(lin, logm_dev, A, K, V, TS, beta_t_dev, beta64_t_dev);
    ctx->launches += 2;
    int flag[2] = {0, 0};
    KCU(cudaMemcpyAsync(flag, d_flag, sizeof(int) * 2, cudaMemcpyDeviceToHost, st));
    KCU(cudaStreamSynchronize(st));
    KCU(cudaGetLastError());
    double ybad[4] = {0, 0, 0, 0};
    if (flag[0]) {   // diagnostics of one word that failed: its largest / smallest count and whether any is not finite
        std::vector<double> col((size_t)n);
        const int vb = word_column >= 0 ? word_column : flag[1];
        for (int r = 0; r < n; ++r)
            cudaMemcpy(&col[r], stats_dev + off[0] + ((size_t)(r / K) * V + vb) * TS + (r % K), sizeof(double), cudaMemcpyDeviceToHost);
        ybad[0] = *std::max_element(col.begin(), col.end());
        ybad[1] = *std::min_element(col.begin(), col.end());
        for (double c : col) if (!std::isfinite(c)) ybad[2] += 1.0;
    }
#undef KCU
    cleanup();
    if (flag[0])
        return fail(ctx, STM_ERR_CUDA, "stm_update_kappa: Newton iteration did not converge for " + std::to_string(flag[0]) +
                                           " word(s), e.g. word " + std::to_string(flag[1]) + " (counts max " +
                                           std::to_string(ybad[0]) + ", min " + std::to_string(ybad[1]) + ", non-finite " +
                                           std::to_string((int)ybad[2]) + ")");
    ctx->kappa_warm = kappa_dev;            // the next call with the same buffer starts from these coefficients
    ctx->kappa_warm_column = word_column;
    return STM_OK;
}

}  // extern "C"
