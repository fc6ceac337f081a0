// spectral.cuh — spectral initialisation of beta on the device.
// Replaces spectral_init / gram / fastAnchor / recover_l2, /root/reference/src/modules/stm.py:30-296
// ("stm.py:N" below).  Included by stm_b200.cu (needs stm_ctx and the CU / CB macros).
//
// Phases (entry points at the bottom):
//   stm_spectral_gram    this rank's documents -> Htilde'Htilde (upper triangle, cuBLAS Dsyrk over dense
//                        document chunks) and diag(Hhat); the ONE buffer a sharded fit all-reduces
//   stm_spectral_finish  Q = gram - Hhat, row-sum check, fastAnchor (K passes over the V' x V' matrix,
//                        all on the device, no host round trip per pass), recover_l2 (one warp per word:
//                        Lawson-Hanson NNLS on the shared K x K normal matrix), back to K x V
//
// Reference behaviour that is kept on purpose (verified against the live reference, oracle/spectral_numpy.py):
//   * Q is NOT row-normalised: stm.py:156 normalises a temporary CSR copy of the CSC matrix;
//   * fastAnchor ranks COLUMN sums of squares but rescales the ROW of that index (stm.py:175, 186, 222);
//   * `basis = np.zeros(K)`: row 0 is never projected and column 0 is never ranked until the last pass
//     (stm.py:176, 216-218, 223);
//   * recover_l2 sees the caller's Q with only the first anchor's row rescaled (stm.py:186, 219);
//   * beta is normalised by its total sum (stm.py:82).
#pragma once

namespace stm_spectral {

constexpr int STRIP = 32;   // rows of Q per CTA in the projection pass
constexpr int CPT = 4;      // columns per thread there (256 threads -> 1024 columns per CTA)

__device__ __forceinline__ double warp_sum_d(double v) {
#pragma unroll
    for (int o = 16; o; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// gram(), per document (stm.py:134-146): wc = sum of kept counts, div = wc (wc - 1), Hhat += c / div.
// A document with fewer than two kept tokens makes the reference's Q NaN (0/0 or inf - inf) and its
// row-sum assertion fail (stm.py:152-154): flagged here.
__global__ void doc_scale_kernel(const long long* __restrict__ doc_ptr, const int* __restrict__ word_id,
                                 const float* __restrict__ count, const int* __restrict__ col_of, long long D,
                                 double* __restrict__ div_out, double* __restrict__ hhat, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long d = warp; d < D; d += nwarps) {
        const long long lo = doc_ptr[d], hi = doc_ptr[d + 1];
        double wc = 0.0;
        for (long long i = lo + lane; i < hi; i += 32)
            if (col_of[word_id[i]] >= 0) wc += (double)count[i];
        wc = warp_sum_d(wc);   // integers: exact in any order
        const double div = wc * (wc - 1.0);
        if (lane == 0) {
            div_out[d] = div;
            if (!(wc >= 2.0)) atomicExch(flag, 1);
        }
        for (long long i = lo + lane; i < hi; i += 32) {
            const int c = col_of[word_id[i]];
            if (c >= 0) atomicAdd(&hhat[c], (double)count[i] / div);
        }
    }
}

// packed upper triangle (row-major, j >= i) of an n x n symmetric matrix
__host__ __device__ inline size_t tri_index(int i, int j, int n) {
    return (size_t)i * n - ((size_t)i * (i - 1)) / 2 + (size_t)(j - i);
}
__host__ __device__ inline size_t tri_size(int n) { return ((size_t)n * (n + 1)) / 2; }

// Htilde'Htilde (stm.py:145-149) accumulated SPARSELY: one warp per document stages the document's kept entries
// (column, count / sqrt(divisor)) in shared memory and adds the m (m + 1) / 2 products of its outer product into
// the packed upper triangle with fp64 reductions (red.global.add.f64, resolved in L2: the 100 MB triangle at
// maxV = 5000 mostly lives there).  Algorithmic work: sum_d m_d^2 / 2 multiply-adds (~0.7 G at BASELINE config 3)
// instead of the 2.5e12 of the densify + Dsyrk version of round 1 (a library GEMM doing ~1000x the sparse work).
__global__ void gram_outer_kernel(const long long* __restrict__ doc_ptr, const int* __restrict__ word_id,
                                  const float* __restrict__ count, const int* __restrict__ col_of,
                                  const double* __restrict__ div, long long D, int n, int cap,
                                  double* __restrict__ tri) {
    extern __shared__ __align__(16) unsigned char sm_raw[];
    const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = blockDim.x >> 5;
    double* hv = reinterpret_cast<double*>(sm_raw) + (size_t)w * cap;
    int* hc = reinterpret_cast<int*>(reinterpret_cast<double*>(sm_raw) + (size_t)nw * cap) + (size_t)w * cap;
    for (long long d = (long long)blockIdx.x * nw + w; d < D; d += (long long)gridDim.x * nw) {
        const long long lo = doc_ptr[d], hi = doc_ptr[d + 1];
        const double s = sqrt(div[d]);
        // compact the kept entries (order within the document is irrelevant for a sum)
        int m = 0;
        for (long long i0 = lo; i0 < hi; i0 += 32) {
            const long long i = i0 + lane;
            const int c = (i < hi) ? col_of[word_id[i]] : -1;
            const unsigned mask = __ballot_sync(0xffffffffu, c >= 0);
            if (c >= 0) {
                const int pos = m + __popc(mask & ((1u << lane) - 1u));
                hc[pos] = c;
                hv[pos] = (double)count[i] / s;
            }
            m += __popc(mask);
        }
        __syncwarp();
        for (int a = 0; a < m; ++a) {
            const int ca = hc[a];
            const double va = hv[a];
            for (int b = lane; b <= a; b += 32) {
                const int cb = hc[b];
                const int i = min(ca, cb), j = max(ca, cb);
                stm::red_add_f64(tri + tri_index(i, j, n), va * hv[b]);
            }
        }
        __syncwarp();
    }
}

// packed upper triangle -> full row-major matrix with the diagonal Hhat subtracted: Q = Htilde'Htilde - Hhat (stm.py:149)
__global__ void expand_tri_kernel(const double* __restrict__ tri, const double* __restrict__ hhat, int n,
                                  double* __restrict__ Q) {
    const long long total = (long long)n * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / n), j = (int)(t % n);
        const double v = tri[(j >= i) ? tri_index(i, j, n) : tri_index(j, i, n)];
        Q[t] = (i == j) ? v - hhat[i] : v;
    }
}

// the assertion at stm.py:152-154: every row sum of Q must be > 0 (NaN fails it too)
__global__ void rowsum_check_kernel(const double* __restrict__ Q, int n, int* __restrict__ flag) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < n; r += nwarps) {
        double s = 0.0;
        for (int c = lane; c < n; c += 32) s += Q[(size_t)r * n + c];
        s = warp_sum_d(s);
        if (lane == 0 && !(s > 0.0)) atomicExch(flag, 2);
    }
}

// ---- fastAnchor (stm.py:160-226) ------------------------------------------------------------------

// column sums of squares over one strip of rows -> part[strip][col]  (first pass only; later passes get
// them from project_kernel)
__global__ void colsq_partial_kernel(const double* __restrict__ Q, int n, double* __restrict__ part) {
    const int r0 = blockIdx.x * STRIP, r1 = min(n, r0 + STRIP);
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * CPT;
    double acc[CPT] = {0.0, 0.0, 0.0, 0.0};
    for (int r = r0; r < r1; ++r)
#pragma unroll
        for (int u = 0; u < CPT; ++u)
            if (c0 + u < n) { const double q = Q[(size_t)r * n + c0 + u]; acc[u] += q * q; }
#pragma unroll
    for (int u = 0; u < CPT; ++u)
        if (c0 + u < n) part[(size_t)blockIdx.x * n + c0 + u] = acc[u];
}

// strips summed in order (deterministic); columns in the zero set of pass `pass` are cleared
// (stm.py:223: row_squared_sum[:, basis] = 0 with the not-yet-filled entries of basis equal to 0)
__global__ void colsq_final_kernel(const double* __restrict__ part, int nstrips, int n, const int* __restrict__ basis,
                                   int pass, double* __restrict__ colsq) {
    const int c = blockIdx.x * blockDim.x + threadIdx.x;
    if (c >= n) return;
    double s = 0.0;
    for (int t = 0; t < nstrips; ++t) s += part[(size_t)t * n + c];
    if (pass > 0) {
        bool zero = (c == 0);
        for (int t = 0; t < pass; ++t) zero |= (basis[t] == c);
        if (zero) s = 0.0;
    }
    colsq[c] = s;
}

// np.argmax (first index of the maximum) -> basis[pass]; one CTA
__global__ void argmax_kernel(const double* __restrict__ colsq, int n, int* __restrict__ basis, int pass) {
    __shared__ double sv[32];
    __shared__ int si[32];
    double bv = -1.0;   // sums of squares are >= 0
    int bi = 0x7fffffff;
    for (int c = threadIdx.x; c < n; c += blockDim.x) {
        const double v = colsq[c];
        if (v > bv || (v == bv && c < bi)) { bv = v; bi = c; }
    }
    for (int o = 16; o; o >>= 1) {
        const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
        const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
        if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
    }
    if ((threadIdx.x & 31) == 0) { sv[threadIdx.x >> 5] = bv; si[threadIdx.x >> 5] = bi; }
    __syncthreads();
    if (threadIdx.x < 32) {
        const int nw = blockDim.x >> 5;
        bv = threadIdx.x < nw ? sv[threadIdx.x] : -1.0;
        bi = threadIdx.x < nw ? si[threadIdx.x] : 0x7fffffff;
        for (int o = 16; o; o >>= 1) {
            const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
            const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
            if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
        }
        if (threadIdx.x == 0) basis[pass] = (bi == 0x7fffffff) ? 0 : bi;   // all-NaN column sums: argmax -> 0
    }
}

// Q[maxind] *= 1 / sqrt(maxval) (stm.py:183-186); the scaled row also goes to rvec and, in the first
// pass, to the caller's copy of Q
__global__ void scale_row_kernel(double* __restrict__ Q, int n, const int* __restrict__ basis, int pass,
                                 const double* __restrict__ colsq, double* __restrict__ rvec, double* __restrict__ Qcaller) {
    const int m = basis[pass];
    const double normalizer = 1.0 / sqrt(colsq[m]);
    for (int c = blockIdx.x * blockDim.x + threadIdx.x; c < n; c += gridDim.x * blockDim.x) {
        const double v = Q[(size_t)m * n + c] * normalizer;
        Q[(size_t)m * n + c] = v;
        rvec[c] = v;
        if (Qcaller) Qcaller[(size_t)m * n + c] = v;
    }
}

// innerproducts = Q @ Q[maxind].T (stm.py:189-194): one warp per row
__global__ void matvec_kernel(const double* __restrict__ Q, const double* __restrict__ rvec, int n,
                              double* __restrict__ ip) {
    const int lane = threadIdx.x & 31;
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int nwarps = (gridDim.x * blockDim.x) >> 5;
    for (int r = warp; r < n; r += nwarps) {
        double s = 0.0;
        for (int c = lane; c < n; c += 32) s += Q[(size_t)r * n + c] * rvec[c];
        s = warp_sum_d(s);
        if (lane == 0) ip[r] = s;
    }
}

// Q -= project with the basis rows of project zeroed (stm.py:206-219), fused with the next pass's column
// sums of squares over ALL rows (stm.py:222).  Zero set of pass i: basis[0..i] and, while basis still has
// unfilled entries (i < K-1), row 0.
__global__ void project_kernel(double* __restrict__ Q, const double* __restrict__ ip, const double* __restrict__ rvec,
                               int n, const int* __restrict__ basis, int pass, int K, double* __restrict__ part) {
    __shared__ unsigned char skip[STRIP];
    __shared__ double ips[STRIP];
    const int r0 = blockIdx.x * STRIP, r1 = min(n, r0 + STRIP);
    if (threadIdx.x < STRIP) {
        const int r = r0 + threadIdx.x;
        bool z = (r == 0 && pass < K - 1);
        for (int t = 0; t <= pass; ++t) z |= (basis[t] == r);
        skip[threadIdx.x] = z;
        ips[threadIdx.x] = r < n ? ip[r] : 0.0;
    }
    __syncthreads();
    const int c0 = (blockIdx.y * blockDim.x + threadIdx.x) * CPT;
    double rv[CPT], acc[CPT];
#pragma unroll
    for (int u = 0; u < CPT; ++u) { rv[u] = (c0 + u < n) ? rvec[c0 + u] : 0.0; acc[u] = 0.0; }
    for (int r = r0; r < r1; ++r) {
        const bool sk = skip[r - r0];
        const double a = ips[r - r0];
#pragma unroll
        for (int u = 0; u < CPT; ++u) {
            if (c0 + u < n) {
                double q = Q[(size_t)r * n + c0 + u];
                if (!sk) { q = q - a * rv[u]; Q[(size_t)r * n + c0 + u] = q; }
                acc[u] += q * q;
            }
        }
    }
#pragma unroll
    for (int u = 0; u < CPT; ++u)
        if (c0 + u < n) part[(size_t)blockIdx.x * n + c0 + u] = acc[u];
}

// ---- recover_l2 (stm.py:229-296) -------------------------------------------------------------------

// M = Q[anchor] (stm.py:245-247)
__global__ void gather_rows_kernel(const double* __restrict__ Qc, int n, const int* __restrict__ basis, int K,
                                   double* __restrict__ M) {
    const long long total = (long long)K * n;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x)
        M[t] = Qc[(size_t)basis[t / n] * n + (t % n)];
}

// One warp per word: min 1/2 w'Pw - q'w, w >= 0 (the reference's QP in w = -x, stm.py:254-285), by the
// Lawson-Hanson active-set method on the normal equations; the passive set's system is solved by a
// Cholesky factorisation held in shared memory.  Anchor words get their one-hot row (stm.py:262-265).
// Shared layout: [P K*K (if p_shared)] then per warp: L K*K | q K | w K | z K | b K | S K ints | state K bytes.
__global__ void nnls_kernel(const double* __restrict__ Pg, const double* __restrict__ QM, const int* __restrict__ basis,
                            int n, int K, int p_shared, double* __restrict__ weights, int* __restrict__ flag) {
    extern __shared__ double sm[];
    const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5, wpb = blockDim.x >> 5;
    const double* P = Pg;
    double* base = sm;
    if (p_shared) {
        for (int t = threadIdx.x; t < K * K; t += blockDim.x) sm[t] = Pg[t];
        P = sm;
        base = sm + K * K;
    }
    __syncthreads();
    const size_t per_warp = (size_t)K * K + 4 * (size_t)K + ((size_t)K * 4 + K + 7) / 8 + 1;
    double* L = base + wib * per_warp;
    double* q = L + (size_t)K * K;
    double* w = q + K;
    double* z = w + K;
    double* b = z + K;
    int* S = reinterpret_cast<int*>(b + K);
    unsigned char* state = reinterpret_cast<unsigned char*>(S + K);   // 0 free, 1 passive, 2 excluded

    // tol = 10 K eps max_j sum_i |P_ij|
    double colmax = 0.0;
    for (int j = lane; j < K; j += 32) {
        double s = 0.0;
        for (int i = 0; i < K; ++i) s += fabs(P[i * K + j]);
        colmax = fmax(colmax, s);
    }
    for (int o = 16; o; o >>= 1) colmax = fmax(colmax, __shfl_xor_sync(0xffffffffu, colmax, o));
    const double tol = 10.0 * K * 2.220446049250313e-16 * colmax;

    for (int word = blockIdx.x * wpb + wib; word < n; word += gridDim.x * wpb) {
        bool is_anchor = false;
        for (int k = lane; k < K; k += 32) is_anchor |= (basis[k] == word);
        if (__any_sync(0xffffffffu, is_anchor)) {
            for (int k = lane; k < K; k += 32) weights[(size_t)word * K + k] = (basis[k] == word) ? 1.0 : 0.0;
            continue;
        }
        for (int k = lane; k < K; k += 32) { q[k] = QM[(size_t)word * K + k]; w[k] = 0.0; state[k] = 0; }
        int ns = 0;
        __syncwarp();
        for (int outer = 0; outer < 3 * K; ++outer) {
            // dual vector g = q - P w over the free indices; first index of its maximum
            double bv = -1e300;
            int bi = 0x7fffffff;
            for (int k = lane; k < K; k += 32) {
                if (state[k] != 0) continue;
                double g = q[k];
                for (int t = 0; t < ns; ++t) g -= P[k * K + S[t]] * w[S[t]];
                if (g > bv) { bv = g; bi = k; }
            }
            for (int o = 16; o; o >>= 1) {
                const double ov = __shfl_xor_sync(0xffffffffu, bv, o);
                const int oi = __shfl_xor_sync(0xffffffffu, bi, o);
                if (ov > bv || (ov == bv && oi < bi)) { bv = ov; bi = oi; }
            }
            if (!(bv > tol) || bi == 0x7fffffff) break;
            const int j = bi;
            // insert j into the ascending passive list
            if (lane == 0) {
                int pos = ns;
                while (pos > 0 && S[pos - 1] > j) { S[pos] = S[pos - 1]; --pos; }
                S[pos] = j;
                state[j] = 1;
            }
            ++ns;
            __syncwarp();
            bool fresh = true;
            while (true) {
                // Cholesky of P[S, S] (column by column, lanes over rows), then two triangular solves
                bool ok = true;
                for (int c = 0; c < ns && ok; ++c) {
                    double dval = 0.0;
                    for (int r = c + lane; r < ns; r += 32) {
                        double s = P[S[r] * K + S[c]];
                        for (int t = 0; t < c; ++t) s -= L[r * ns + t] * L[c * ns + t];
                        b[r] = s;   // staging
                        if (r == c) dval = s;
                    }
                    dval = __shfl_sync(0xffffffffu, dval, 0);
                    if (!(dval > 0.0)) { ok = false; break; }
                    const double lcc = sqrt(dval);
                    __syncwarp();
                    for (int r = c + lane; r < ns; r += 32) L[r * ns + c] = (r == c) ? lcc : b[r] / lcc;
                    __syncwarp();
                }
                int jpos = 0;
                if (ok) {
                    for (int r = lane; r < ns; r += 32) b[r] = q[S[r]];
                    __syncwarp();
                    for (int c = 0; c < ns; ++c) {          // L y = q_S (y overwrites b)
                        if (lane == 0) b[c] = b[c] / L[c * ns + c];
                        __syncwarp();
                        const double yc = b[c];
                        for (int r = c + 1 + lane; r < ns; r += 32) b[r] -= L[r * ns + c] * yc;
                        __syncwarp();
                    }
                    for (int c = ns - 1; c >= 0; --c) {      // L' z = y
                        if (lane == 0) z[c] = b[c] / L[c * ns + c];
                        __syncwarp();
                        const double zc = z[c];
                        for (int r = lane; r < c; r += 32) b[r] -= L[c * ns + r] * zc;
                        __syncwarp();
                    }
                    for (int t = 0; t < ns; ++t) if (S[t] == j) jpos = t;
                }
                if (!ok || (fresh && z[jpos] <= 0.0)) {
                    if (!fresh) { if (lane == 0) atomicExch(flag, 3); }
                    // rounding guard of Lawson & Hanson: drop the fresh index and do not pick it again
                    // until the solution has moved
                    if (lane == 0) {
                        int pos = 0;
                        for (int t = 0; t < ns; ++t) if (S[t] != j) S[pos++] = S[t];
                        state[j] = 2;
                        w[j] = 0.0;
                    }
                    --ns;
                    __syncwarp();
                    break;
                }
                fresh = false;
                bool allpos = true;
                for (int t = lane; t < ns; t += 32) allpos &= (z[t] > 0.0);
                allpos = __all_sync(0xffffffffu, allpos);
                if (allpos) {
                    for (int t = lane; t < ns; t += 32) w[S[t]] = z[t];
                    for (int k = lane; k < K; k += 32) if (state[k] == 2) state[k] = 0;
                    __syncwarp();
                    break;
                }
                // step towards z until the first passive coordinate hits zero
                double av = 1e300;
                int ai = 0x7fffffff;
                for (int t = lane; t < ns; t += 32) {
                    if (z[t] <= 0.0) {
                        const double ws = w[S[t]];
                        const double r = ws / (ws - z[t]);
                        if (r < av) { av = r; ai = t; }
                    }
                }
                for (int o = 16; o; o >>= 1) {
                    const double ov = __shfl_xor_sync(0xffffffffu, av, o);
                    const int oi = __shfl_xor_sync(0xffffffffu, ai, o);
                    if (ov < av || (ov == av && oi < ai)) { av = ov; ai = oi; }
                }
                for (int t = lane; t < ns; t += 32) {
                    const double ws = w[S[t]];
                    double nv = ws + av * (z[t] - ws);
                    if (t == ai) nv = 0.0;
                    w[S[t]] = nv;
                }
                __syncwarp();
                int pos = 0;
                if (lane == 0) {
                    for (int t = 0; t < ns; ++t) {
                        const int k = S[t];
                        if (w[k] <= 0.0) { w[k] = 0.0; state[k] = 0; }
                        else S[pos++] = k;
                    }
                }
                ns = __shfl_sync(0xffffffffu, pos, 0);
                __syncwarp();
                if (ns == 0) break;
            }
        }
        __syncwarp();
        for (int k = lane; k < K; k += 32) weights[(size_t)word * K + k] = w[k];
        __syncwarp();
    }
}

// A = weights.T * wprob; column sums over words (stm.py:290-292), one CTA per topic, fixed-order tree
__global__ void topic_sum_kernel(const double* __restrict__ weights, const double* __restrict__ wprob, int n, int K,
                                 double* __restrict__ tsum) {
    __shared__ double s[256];
    const int k = blockIdx.x;
    double acc = 0.0;
    for (int i = threadIdx.x; i < n; i += 256) acc += weights[(size_t)i * K + k] * wprob[i];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) tsum[k] = s[0];
}

__global__ void fill_d_kernel(double* __restrict__ p, long long n, double v) {
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        p[t] = v;
}

// beta_new[:, keep] = A / rowsum(A); beta_new += 0.001 / V (stm.py:78-81); beta_new was pre-filled with 0
__global__ void beta_scatter_kernel(const double* __restrict__ weights, const double* __restrict__ wprob,
                                    const double* __restrict__ tsum, const int* __restrict__ keep, int n, int K, int V,
                                    double* __restrict__ beta) {
    const long long total = (long long)n * K;
    const double eps = 0.001 / V;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < total;
         t += (long long)gridDim.x * blockDim.x) {
        const int i = (int)(t / K), k = (int)(t % K);
        beta[(size_t)k * V + keep[i]] = (weights[t] * wprob[i]) / tsum[k] + eps;
    }
}

// total sum in two fixed-order stages, then beta /= total (stm.py:82)
__global__ void total_partial_kernel(const double* __restrict__ x, long long n, double* __restrict__ part) {
    __shared__ double s[256];
    double acc = 0.0;
    for (long long t = (long long)blockIdx.x * 256 + threadIdx.x; t < n; t += (long long)gridDim.x * 256) acc += x[t];
    s[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o; o >>= 1) {
        if (threadIdx.x < o) s[threadIdx.x] += s[threadIdx.x + o];
        __syncthreads();
    }
    if (threadIdx.x == 0) part[blockIdx.x] = s[0];
}

__global__ void total_scale_kernel(double* __restrict__ x, long long n, const double* __restrict__ part, int nparts) {
    __shared__ double tot;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int t = 0; t < nparts; ++t) s += part[t];
        tot = s;
    }
    __syncthreads();
    const double d = tot;
    for (long long t = (long long)blockIdx.x * blockDim.x + threadIdx.x; t < n; t += (long long)gridDim.x * blockDim.x)
        x[t] = x[t] / d;
}

}  // namespace stm_spectral

extern "C" {

int stm_spectral_gram(stm_ctx* ctx, int n_keep, const int32_t* keep, double* gram_dev, void* stream) {
    using namespace stm_spectral;
    if (!ctx) return STM_ERR_INVALID;
    if (!ctx->d_doc_ptr) return fail(ctx, STM_ERR_NO_CORPUS, "stm_spectral_gram: stm_set_corpus has not been called");
    if (n_keep < 1 || n_keep > ctx->V || !keep || !gram_dev)
        return fail(ctx, STM_ERR_INVALID, "stm_spectral_gram: bad keep list or NULL output");
    STM_ON_DEVICE(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    const int n = n_keep;
    std::vector<int> col_of(ctx->V, -1);
    for (int i = 0; i < n; ++i) {
        if (keep[i] < 0 || keep[i] >= ctx->V || col_of[keep[i]] != -1)
            return fail(ctx, STM_ERR_INVALID, "stm_spectral_gram: keep must hold distinct word ids in [0, V)");
        col_of[keep[i]] = i;
    }
    const int64_t D = ctx->D;
    int* d_col = nullptr; double* d_div = nullptr; int* d_flag = nullptr;
    auto cleanup = [&]() { cudaFree(d_col); cudaFree(d_div); cudaFree(d_flag); };
#define SCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
        }                                                                                            \
    } while (0)
    SCU(cudaMalloc(&d_col, sizeof(int) * ctx->V));
    SCU(cudaMalloc(&d_div, sizeof(double) * std::max<int64_t>(D, 1)));
    SCU(cudaMalloc(&d_flag, sizeof(int)));
    SCU(cudaMemcpyAsync(d_col, col_of.data(), sizeof(int) * ctx->V, cudaMemcpyHostToDevice, st));
    SCU(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    SCU(cudaMemsetAsync(gram_dev, 0, sizeof(double) * ((size_t)n * n + n), st));
    double* hhat = gram_dev + (size_t)n * n;
    if (D > 0) {
        doc_scale_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(ctx->d_doc_ptr, ctx->d_word_id, ctx->d_count, d_col, D,
                                                            d_div, hhat, d_flag);
        // one warp per document; shared memory: 12 bytes per staged entry, capacity = the longest document
        const int cap = (std::max(ctx->n_max, 1) + 1) & ~1;
        int warps = (int)std::min<size_t>(8, (size_t)(ctx->max_smem - 1024) / ((size_t)cap * 12));
        if (warps < 1) { cleanup(); return fail(ctx, STM_ERR_UNSUPPORTED, "stm_spectral_gram: document too long for the staging buffer"); }
        const size_t smem = (size_t)warps * cap * 12;
        SCU(cudaFuncSetAttribute(gram_outer_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int per_sm = std::max<int>(1, std::min<int>(8, (int)((size_t)ctx->max_smem / std::max<size_t>(smem, 1))));
        gram_outer_kernel<<<ctx->sm_count * per_sm, warps * 32, smem, st>>>(ctx->d_doc_ptr, ctx->d_word_id, ctx->d_count,
                                                                            d_col, d_div, (long long)D, n, cap, gram_dev);
        ctx->launches += 2;
        SCU(cudaGetLastError());
    }
    int flag = 0;
    SCU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));
#undef SCU
    cleanup();
    if (flag)
        return fail(ctx, STM_ERR_INVALID,
                    "Encountered zeroes in Q row sums, can not normalize. (a document has fewer than two kept tokens)");
    return STM_OK;
}

int stm_spectral_finish(stm_ctx* ctx, int n_keep, const int32_t* keep, const double* wprob_keep, double* gram_dev,
                        double* beta_kv_dev, int32_t* anchor_out, void* stream) {
    using namespace stm_spectral;
    if (!ctx) return STM_ERR_INVALID;
    const int n = n_keep, K = ctx->K, V = ctx->V;
    if (n < 1 || n > V || !keep || !wprob_keep || !gram_dev || !beta_kv_dev)
        return fail(ctx, STM_ERR_INVALID, "stm_spectral_finish: bad arguments");
    if (K > n) return fail(ctx, STM_ERR_INVALID, "stm_spectral_finish: more topics than kept words");
    STM_ON_DEVICE(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    const int nstrips = (n + STRIP - 1) / STRIP;
    const dim3 sgrid(nstrips, (n + 256 * CPT - 1) / (256 * CPT));
    double *Q = nullptr, *part = nullptr, *small = nullptr, *M = nullptr, *QM = nullptr, *W = nullptr;
    int *d_basis = nullptr, *d_keep = nullptr, *d_flag = nullptr;
    auto cleanup = [&]() {
        cudaFree(Q); cudaFree(part); cudaFree(small); cudaFree(M); cudaFree(QM); cudaFree(W);
        cudaFree(d_basis); cudaFree(d_keep); cudaFree(d_flag);
    };
#define SCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
        }                                                                                            \
    } while (0)
#define SCB(call)                                                                                    \
    do {                                                                                             \
        if ((call) != CUBLAS_STATUS_SUCCESS) {                                                       \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + " failed");                          \
        }                                                                                            \
    } while (0)
    const int nparts = 1024;
    // small: colsq n | rvec n | ip n | wprob n | P K*K | tsum K | total partials
    const size_t small_len = 4 * (size_t)n + (size_t)K * K + K + nparts;
    SCU(cudaMalloc(&Q, sizeof(double) * (size_t)n * n));              // fastAnchor's working copy
    SCU(cudaMalloc(&part, sizeof(double) * (size_t)nstrips * n));
    SCU(cudaMalloc(&small, sizeof(double) * small_len));
    SCU(cudaMalloc(&M, sizeof(double) * (size_t)K * n));
    SCU(cudaMalloc(&QM, sizeof(double) * (size_t)n * K));
    SCU(cudaMalloc(&W, sizeof(double) * (size_t)n * K));
    SCU(cudaMalloc(&d_basis, sizeof(int) * K));
    SCU(cudaMalloc(&d_keep, sizeof(int) * n));
    SCU(cudaMalloc(&d_flag, sizeof(int)));
    double *colsq = small, *rvec = small + n, *ip = small + 2 * (size_t)n, *wprob = small + 3 * (size_t)n,
           *P = small + 4 * (size_t)n, *tsum = P + (size_t)K * K, *tpart = tsum + K;
    SCU(cudaMemsetAsync(d_flag, 0, sizeof(int), st));
    SCU(cudaMemsetAsync(d_basis, 0, sizeof(int) * K, st));
    SCU(cudaMemcpyAsync(d_keep, keep, sizeof(int) * n, cudaMemcpyHostToDevice, st));
    SCU(cudaMemcpyAsync(wprob, wprob_keep, sizeof(double) * n, cudaMemcpyHostToDevice, st));

    // ---- Q = Htilde'Htilde - Hhat, row-sum assertion (stm.py:149-154) ----
    // gram_dev holds the packed upper triangle of Htilde'Htilde (what a sharded fit all-reduced) and diag(Hhat):
    // expand it to the full matrix in place, through a copy of the triangle (fastAnchor's working buffer, not yet in use)
    double* Qc = gram_dev;   // becomes the caller's Q
    SCU(cudaMemcpyAsync(Q, gram_dev, sizeof(double) * tri_size(n), cudaMemcpyDeviceToDevice, st));   // Q: scratch until fastAnchor
    expand_tri_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(Q, gram_dev + (size_t)n * n, n, Qc);
    rowsum_check_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(Qc, n, d_flag);
    ctx->launches += 2;
    int flag = 0;
    SCU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));
    if (flag) {
        cleanup();
        return fail(ctx, STM_ERR_INVALID, "Encountered zeroes in Q row sums, can not normalize.");
    }
    SCU(cudaMemcpyAsync(Q, Qc, sizeof(double) * (size_t)n * n, cudaMemcpyDeviceToDevice, st));

    // ---- fastAnchor (stm.py:160-226): K passes, no host synchronisation ----
    colsq_partial_kernel<<<sgrid, 256, 0, st>>>(Q, n, part);
    ctx->launches++;
    for (int pass = 0; pass < K; ++pass) {
        colsq_final_kernel<<<(n + 255) / 256, 256, 0, st>>>(part, nstrips, n, d_basis, pass, colsq);
        argmax_kernel<<<1, 1024, 0, st>>>(colsq, n, d_basis, pass);
        scale_row_kernel<<<(n + 255) / 256, 256, 0, st>>>(Q, n, d_basis, pass, colsq, rvec, pass == 0 ? Qc : nullptr);
        matvec_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(Q, rvec, n, ip);
        project_kernel<<<sgrid, 256, 0, st>>>(Q, ip, rvec, n, d_basis, pass, K, part);
        ctx->launches += 5;
    }

    // ---- recover_l2 (stm.py:229-296) ----
    gather_rows_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(Qc, n, d_basis, K, M);
    ctx->launches++;
    SCB(cublasSetStream(ctx->cublas, st));
    const double one = 1.0, zero = 0.0;
    // row-major M (K x n) is the column-major n x K matrix Mc; P = Mc' Mc; (Q M')' = Mc' Qcc with Qcc = Qc'
    SCB(cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, K, K, n, &one, M, n, M, n, &zero, P, K));
    SCB(cublasDgemm(ctx->cublas, CUBLAS_OP_T, CUBLAS_OP_N, K, n, n, &one, M, n, Qc, n, &zero, QM, K));
    {
        const size_t per_warp = ((size_t)K * K + 4 * (size_t)K + ((size_t)K * 4 + K + 7) / 8 + 1) * sizeof(double);
        const size_t pbytes = (size_t)K * K * sizeof(double);
        const size_t budget = (size_t)ctx->max_smem - 1024;
        int p_shared = (pbytes + per_warp <= budget) ? 1 : 0;
        if (!p_shared && per_warp > budget) {
            cleanup();
            return fail(ctx, STM_ERR_UNSUPPORTED, "stm_spectral_finish: K too large for the NNLS kernel");
        }
        int wpb = (int)std::min<size_t>(8, (budget - (p_shared ? pbytes : 0)) / per_warp);
        const size_t smem = (p_shared ? pbytes : 0) + wpb * per_warp;
        SCU(cudaFuncSetAttribute(nnls_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        const int grid = std::min((n + wpb - 1) / wpb, ctx->sm_count * 4);
        nnls_kernel<<<grid, wpb * 32, smem, st>>>(P, QM, d_basis, n, K, p_shared, W, d_flag);
        ctx->launches++;
    }
    topic_sum_kernel<<<K, 256, 0, st>>>(W, wprob, n, K, tsum);
    fill_d_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(beta_kv_dev, (long long)K * V, 0.001 / V);
    beta_scatter_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(W, wprob, tsum, d_keep, n, K, V, beta_kv_dev);
    total_partial_kernel<<<nparts, 256, 0, st>>>(beta_kv_dev, (long long)K * V, tpart);
    total_scale_kernel<<<ctx->sm_count * 4, 256, 0, st>>>(beta_kv_dev, (long long)K * V, tpart, nparts);
    ctx->launches += 5;
    std::vector<int> basis(K, 0);
    SCU(cudaMemcpyAsync(&flag, d_flag, sizeof(int), cudaMemcpyDeviceToHost, st));
    SCU(cudaMemcpyAsync(basis.data(), d_basis, sizeof(int) * K, cudaMemcpyDeviceToHost, st));
    SCU(cudaStreamSynchronize(st));
    SCU(cudaGetLastError());
#undef SCU
#undef SCB
    cleanup();
    if (flag == 3) return fail(ctx, STM_ERR_NOT_PD, "stm_spectral_finish: anchor rows are linearly dependent (singular QP)");
    if (anchor_out) for (int k = 0; k < K; ++k) anchor_out[k] = basis[k];
    return STM_OK;
}

}  // extern "C"
