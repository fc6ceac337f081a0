// bfgs_slots.cuh — kernel A, second generation: document SLOTS (included after estep_kernel.cuh).
//
// Same arithmetic as stm::bfgs_kernel (stm.py:536-545 -> scipy.optimize.minimize(method="BFGS"),
// scipy/optimize/_optimize.py:1345-1526, _linesearch.py, _dcsrch.py), different mapping.  Round 1 measured that
// ~75 % of kernel A's warp instructions were the warp-uniform scalar code around the K x n_d contraction (line-search
// state machine, logsumexp, accept bookkeeping), executed redundantly by 32 lanes, and that instruction issue / fetch
// is the limiter.  Here a warp owns up to SLOT_GMAX document SLOTS (tile resident in tensor memory or shared memory):
//   * objective evaluations stay warp-cooperative (lane <-> word in the contraction, lane <-> topic for the K-vectors)
//     and run one slot after the other;
//   * the scalar part — logsumexp, f, SciPy's dcsrch / Wolfe-2 / zoom state machine, BFGS bookkeeping — runs ONCE per
//     round with lane g <-> slot g, its state in registers of that lane;
//   * the K-vectors of a slot live in shared memory (indexed by slot, so the evaluation code exists once);
//   * results travel between the two through a per-slot mailbox (SlotBox) in shared memory.
// A finished slot is refilled from the work queue independently of its neighbours (no lock step between documents).
#pragma once

#ifndef STM_SLOTS_MAX_THREADS
#define STM_SLOTS_MAX_THREADS 256   // launch bound (register budget 65536 / this)
#endif

#ifndef STM_SLOTS_TIMING
#define STM_SLOTS_TIMING 0   // variant builds: clock64() per phase into P.dbg_cycles (tools/gpu_perf.py prints them)
#endif
#if STM_SLOTS_TIMING
#define SLT_T(var) const long long var = clock64()
#define SLT_ACC(slot, t0, t1) dbg_t[slot] += (t1) - (t0)
#define SLT_CNT(slot) dbg_t[slot] += 1
#else
#define SLT_T(var)
#define SLT_ACC(slot, t0, t1)
#define SLT_CNT(slot)
#endif

namespace stm {

constexpr int SLOT_GMAX = 4;

// per-slot mailbox: written by one lane, read through a warp-uniform address by all
struct SlotBox {
    double alpha;      // trial step of the next evaluation (lane g -> evaluation)
    double f[2];       // objective at the two memoised points
    double scale[2];   // N_d / sum_k exp(eta_k) at the two memoised points (gradient factor, stm.py:955-957)
    double res[8];     // evaluation -> lane g: m, cnt, ssum, quad2, logprod, (S d - a).p, ex.p, dphi (memo hit)
    double acc[4];     // accept / init -> lane g: gnorm, derphi0, |g|^2
    double Nsum;       // np.sum(word_count)
    int cur;           // memo buffer holding the most recent point
    int hc[2];         // memo buffer valid
    int kind;          // evaluation outcome: 0 fresh, 1 memo hit
    int code;          // accept outcome: 0 continue, 1 converged / stalled, 2 non-finite objective
    int k_it;          // BFGS iterations done (incl. the one being accepted)
    int d, n;          // document index, distinct words
};
static_assert(sizeof(SlotBox) % 16 == 0, "SlotBox must keep the K-vectors 16-byte aligned");

// bytes of one slot's small block: mailbox + 7 K-vectors (x, p, mu, a, g, xt[2]) + ex[2] + counts (+ own a_k scratch)
__host__ __device__ inline size_t slots_small_bytes(int K, int TS, int n_cap) {
    const int K1 = K - 1, KS = (K1 + 1) & ~1;
    size_t b = sizeof(SlotBox) + (size_t)(7 * KS + 2 * TS) * 8 + (size_t)((n_cap + 3) & ~3) * 4;
    if (n_cap > 3 * KS + 2 * TS) b += (size_t)n_cap * 8;
    return (b + 127) & ~(size_t)127;
}
__host__ __device__ inline size_t slots_tile_bytes(int n_cap, int TS) {
    return ((size_t)n_cap * TS * 4 + 127) & ~(size_t)127;
}

struct SlotView {
    SlotBox* box;
    double *x, *p, *mu, *a, *g, *xt, *ex;   // xt: [2][KS], ex: [2][TS]
    float* cw;
    double* wv;       // [n_cap] scratch of the a_k precompute (may alias g..ex)
    float* tile;      // shared-memory tile (nullptr for a TMEM slot)
    uint32_t taddr;   // TMEM address of the slot's first column
    bool is_tm;
};

template <int KPL, int J>
__global__ void __launch_bounds__(STM_SLOTS_MAX_THREADS, 1) bfgs_slots_kernel(const EstepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = __reduce_max_sync(STM_FULL, (int)(threadIdx.x >> 5));   // provably uniform
    const int W = blockDim.x >> 5;
    const int K = P.K, K1 = K - 1, TS = P.TS;
    const int KS = (K1 + 1) & ~1;
    const int TW = P.tm_slots;       // TMEM slots per warp
    const int NS = P.smem_tiles;     // shared-memory tiles per CTA (dealt to the warps round robin)
    const int CS = (K + 1) & ~1;     // TMEM columns per word slot
    int G = TW;
    for (int s = warp; s < NS; s += W) ++G;
    const int n_tm = W * TW;
    const size_t small_b = (size_t)P.smem_small;
    const size_t tile_b = slots_tile_bytes(P.n_cap, TS);
    unsigned char* tiles0 = smem_raw + (size_t)(n_tm + NS) * small_b;
    const bool own_wv = P.n_cap > 3 * KS + 2 * TS;

    __shared__ uint32_t tm_base_s;
    __shared__ uint64_t mbar_s[STM_SLOTS_MAX_THREADS / 32];
    uint64_t* mbar = &mbar_s[warp];
    if (lane == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;
    if (TW > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_base_s)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const uint32_t tm_quarter = (TW > 0) ? tm_base_s + ((uint32_t)((warp & 3) * 32) << 16) : 0u;

    auto view = [&](int g) -> SlotView {
        SlotView v;
        const int sid = (g < TW) ? warp * TW + g : n_tm + warp + (g - TW) * W;
        unsigned char* b = smem_raw + (size_t)sid * small_b;
        v.box = reinterpret_cast<SlotBox*>(b);
        double* vec = reinterpret_cast<double*>(b + sizeof(SlotBox));
        v.x = vec; v.p = vec + KS; v.mu = vec + 2 * KS; v.a = vec + 3 * KS; v.g = vec + 4 * KS;
        v.xt = vec + 5 * KS; v.ex = vec + 7 * KS;
        v.cw = reinterpret_cast<float*>(vec + 7 * KS + 2 * TS);
        v.wv = own_wv ? reinterpret_cast<double*>(v.cw + ((P.n_cap + 3) & ~3)) : v.g;
        v.is_tm = g < TW;
        v.tile = v.is_tm ? nullptr : reinterpret_cast<float*>(tiles0 + (size_t)(warp + (g - TW) * W) * tile_b);
        v.taddr = v.is_tm ? tm_quarter + (uint32_t)(((warp >> 2) * TW + g) * P.tm_cols) : 0u;
        return v;
    };

    // zero the K-vector blocks once: the pads of ex (k in [K, TS)) must stay zero for the contraction
    for (int g = 0; g < G; ++g) {
        SlotView v = view(g);
        for (int i = lane; i < 7 * KS + 2 * TS; i += 32) v.x[i] = 0.0;
    }
    __syncwarp();

    // diagonal of siginv, lane-distributed
    double Sd[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        Sd[i] = (k < K1) ? P.prior[k] : 0.0;
    }
    const int gslot0 = (blockIdx.x * W + warp) * SLOT_GMAX;   // global slot index -> BFGS inverse Hessian scratch
    const int maxiter = K1 * 200;
    const double c1 = 1e-4, c2 = 0.9, xtol = 1e-14, stpmin = 1e-100, stpmax = 1e100;

    // ---- per-lane state of slot `lane` (lanes >= G idle in the scalar phases) -------------------------------------
    int st = 0;              // 0 empty, 1 active, 2 finished (to be written out), 3 closed (queue drained)
    int ls = LS_INIT, warnflag = 0, nfev = 0, k_it = 0;
    double alpha = 0.0, f_eval = 0.0, dphi = 0.0, Nsum = 0.0, Nint = 0.0;
    LsState S;
    S.old_fval = 0.0; S.old_old_fval = 0.0; S.gnorm = 0.0; S.derphi0 = 0.0; S.f2 = 0.0;
    S.finit = 0.0; S.ginit = 0.0; S.gtest = 0.0; S.width = 0.0; S.width1 = 0.0; S.stx = 0.0; S.fx = 0.0; S.gx = 0.0;
    S.sty = 0.0; S.fy = 0.0; S.gy = 0.0; S.stmin = 0.0; S.stmax = 0.0;
    S.alpha0 = 0.0; S.phi_a0 = 0.0; S.derphi_a0 = 0.0; S.a_lo = 0.0; S.a_hi = 0.0; S.phi_lo = 0.0; S.phi_hi = 0.0;
    S.derphi_lo = 0.0; S.phi_rec = 0.0; S.a_rec = 0.0;
    S.brackt = 0; S.stage = 1; S.w1_it = 0; S.w2_i = 0; S.z_i = 0; S.pad_ = 0;
    if (lane >= G) st = 3;
    bool drained = false;    // uniform: the work queue is empty
#if STM_SLOTS_TIMING
    long long dbg_t[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    for (;;) {
        SLT_T(t_r0);
        SLT_CNT(8);
        // =================== refill: empty slots pull the next document ============================================
        for (unsigned m = __ballot_sync(STM_FULL, st == 0); m != 0u; m &= m - 1u) {
            const int g = __ffs((int)m) - 1;
            int qi = P.n_docs;
            if (!drained) {
                qi = 0;
                if (lane == 0) qi = (int)atomicAdd(P.queue, 1u);
                qi = __reduce_max_sync(STM_FULL, qi);
            }
            if (qi >= P.n_docs) {
                drained = true;
                if (lane == g) st = 3;
                continue;
            }
            const SlotView v = view(g);
            const int d = P.docs[qi];
            const long long p0 = P.doc_ptr[d];
            const int n = (int)(P.doc_ptr[d + 1] - p0);
            const int asp = P.aspect ? P.aspect[d] : 0;
            const float* beta_a = P.beta_t + (size_t)asp * P.V * TS;
            double a[KPL];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                a[i] = 0.0;
                if (k < K1) {
                    v.x[k] = P.eta[(size_t)d * K1 + k];
                    v.mu[k] = P.mu[(size_t)d * K1 + k];
                    v.p[k] = 0.0;
                }
            }
            double nsum_l = 0.0;
            if (v.is_tm) {
                // ---- TMEM slot: global -> registers -> tcgen05.st, column sums on the way ----------------------
                int widr[J];
                double rv[J];
                const float4* rowp[J];
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int w = lane + 32 * j;
                    const bool ok = w < n;
                    widr[j] = ok ? P.word_id[p0 + w] : 0;
                    const float c = ok ? P.count[p0 + w] : 0.f;
                    if (w < P.n_cap) v.cw[w] = c;
                    nsum_l += (double)c;
                    rowp[j] = reinterpret_cast<const float4*>(beta_a + (size_t)widr[j] * TS);
                    rv[j] = 0.0;
                }
                int c0 = 0;
                for (; c0 + 8 <= CS; c0 += 8) {
                    float4 b0[J], b1[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) { b0[j] = __ldg(rowp[j] + (c0 >> 2)); b1[j] = __ldg(rowp[j] + (c0 >> 2) + 1); }
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        rv[j] += beta_f2d(b0[j].x); rv[j] += beta_f2d(b0[j].y); rv[j] += beta_f2d(b0[j].z); rv[j] += beta_f2d(b0[j].w);
                        rv[j] += beta_f2d(b1[j].x); rv[j] += beta_f2d(b1[j].y); rv[j] += beta_f2d(b1[j].z); rv[j] += beta_f2d(b1[j].w);
                        tm_st8(v.taddr + (uint32_t)(j * CS + c0), b0[j], b1[j]);
                    }
                }
                if ((CS - c0) & 4) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const float4 b = __ldg(rowp[j] + (c0 >> 2));
                        rv[j] += beta_f2d(b.x); rv[j] += beta_f2d(b.y); rv[j] += beta_f2d(b.z); rv[j] += beta_f2d(b.w);
                        tm_st4(v.taddr + (uint32_t)(j * CS + c0), b);
                    }
                    c0 += 4;
                }
                if ((CS - c0) & 2) {
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        const float2 b = __ldg(reinterpret_cast<const float2*>(rowp[j]) + (c0 >> 1));
                        rv[j] += beta_f2d(b.x); rv[j] += beta_f2d(b.y);
                        tm_st2(v.taddr + (uint32_t)(j * CS + c0), b);
                    }
                }
                tm_wait_st();
                __syncwarp();
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int w = lane + 32 * j;
                    rv[j] = (double)((w < P.n_cap) ? v.cw[w] : 0.f) / rv[j];
                }
                // a_k = sum_v beta_kv c_v / colsum_v (stm.py:954): rows re-read from L2, lane <-> topic
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int cnt = min(32, n - 32 * j);
#pragma unroll 4
                    for (int l = 0; l < cnt; ++l) {
                        const int w = __shfl_sync(STM_FULL, widr[j], l);
                        const double r = __shfl_sync(STM_FULL, rv[j], l);
                        const float* row = beta_a + (size_t)w * TS;
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < K) a[i] += beta_f2d(__ldg(row + k)) * r;
                        }
                    }
                }
            } else {
                // ---- shared-memory slot: stage counts, TMA-gather the beta rows ------------------------------
                fence_proxy_async();
                __syncwarp();
                if (lane == 0) mbar_expect_tx(mbar, (uint32_t)(n * TS * 4));
                __syncwarp();
                for (int w = lane; w < n; w += 32) {
                    const int wid = P.word_id[p0 + w];
                    const float c = P.count[p0 + w];
                    v.cw[w] = c;
                    nsum_l += (double)c;
                    tma_row_g2s(v.tile + (size_t)w * TS, beta_a + (size_t)wid * TS, (uint32_t)(TS * 4), mbar);
                }
                mbar_wait(mbar, parity);
                parity ^= 1;
                __syncwarp();
                for (int w = lane; w < n; w += 32) {
                    const float4* row = reinterpret_cast<const float4*>(v.tile + (size_t)w * TS);
                    double cs = 0.0;
                    for (int q = 0; q < TS / 4; ++q) {
                        const float4 b = row[q];
                        cs += beta_f2d(b.x); cs += beta_f2d(b.y); cs += beta_f2d(b.z); cs += beta_f2d(b.w);
                    }
                    v.wv[w] = (double)v.cw[w] / cs;
                }
                __syncwarp();
                for (int w = 0; w < n; ++w) {
                    const double r = v.wv[w];
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int k = lane + 32 * i;
                        if (k < K) a[i] += beta_f2d(v.tile[(size_t)w * TS + k]) * r;
                    }
                }
                __syncwarp();
                if (!own_wv) for (int i = lane; i < 3 * KS + 2 * TS; i += 32) v.g[i] = 0.0;   // wv aliased g, xt, ex
            }
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                if (k < K1) v.a[k] = a[i];
            }
            const double Ns = warp_sum(nsum_l);            // np.sum(word_count)       stm.py:955
            if (lane == 0) {
                v.box->alpha = 0.0; v.box->cur = 0; v.box->hc[0] = 0; v.box->hc[1] = 0; v.box->k_it = 0;
                v.box->d = d; v.box->n = n; v.box->Nsum = Ns;
            }
            if (lane == g) {
                st = 1; ls = LS_INIT; warnflag = 0; nfev = 0; k_it = 0; alpha = 0.0;
                Nsum = Ns;
                Nint = (double)(long long)Ns;              // int(np.sum(word_count))  stm.py:933
                S.old_fval = 0.0; S.old_old_fval = 0.0; S.gnorm = 0.0; S.derphi0 = 0.0;
                S.brackt = 0; S.stage = 1; S.w1_it = 0; S.w2_i = 0; S.z_i = 0;
            }
            __syncwarp();
        }
        const unsigned active = __ballot_sync(STM_FULL, st == 1);
        if (active == 0u) break;
        SLT_T(t_r1);
        SLT_ACC(0, t_r0, t_r1);

        // =================== evaluations: f, g.p at x + alpha p, one slot after the other ==========================
        for (unsigned m = active; m != 0u; m &= m - 1u) {
            const int g = __ffs((int)m) - 1;
            const SlotView v = view(g);
            const double al = v.box->alpha;
            const int cur = v.box->cur;
            const int n = v.box->n;
            const double* xt_c = v.xt + cur * KS;
            const double* xt_o = v.xt + (cur ^ 1) * KS;
            double xn[KPL], pk[KPL], muk[KPL], ak[KPL];
            bool same0 = v.box->hc[cur] != 0, same1 = v.box->hc[cur ^ 1] != 0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                xn[i] = 0.0; pk[i] = 0.0; muk[i] = 0.0; ak[i] = 0.0;
                if (k < K1) {
                    pk[i] = v.p[k]; muk[i] = v.mu[k]; ak[i] = v.a[k];
                    xn[i] = __dadd_rn(v.x[k], __dmul_rn(al, pk[i]));
                    same0 = same0 && (xn[i] == xt_c[k]);
                    same1 = same1 && (xn[i] == xt_o[k]);
                }
            }
            // Memoisation.  SciPy's ScalarFunction re-uses f and g for a trial point identical to the LAST one
            // (scipy/_differentiable_functions.py:391-401); f is deterministic, so re-using the last TWO points is
            // value-identical and removes the A,B,A,B,... tail of a collapsing dcsrch interval.
            const bool hit0 = __all_sync(STM_FULL, same0);
            const bool hit1 = !hit0 && __all_sync(STM_FULL, same1);
            if (hit0 || hit1) {
                const int sel = hit0 ? cur : (cur ^ 1);
                const double sc = v.box->scale[sel];
                const double* xs = v.xt + sel * KS;
                const double* es = v.ex + sel * TS;
                double dp_l = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = lane + 32 * i;
                    if (k < K1) {
                        const double dk = xs[k] - muk[i];
                        const double gt = Sd[i] * dk - (ak[i] - sc * es[k]);
                        dp_l += gt * pk[i];
                    }
                }
                const double dp = warp_sum(dp_l);
                if (lane == 0) { v.box->cur = sel; v.box->kind = 1; v.box->res[7] = dp; }
                SLT_CNT(10);
                continue;
            }
            SLT_CNT(9);
            SLT_T(t_e0);
            const int nb = cur ^ 1;
            double* xt_n = v.xt + nb * KS;
            double* v0 = v.ex + nb * TS;
            double et[KPL], ex[KPL];
            double mx = -INFINITY;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                if (k < K1) xt_n[k] = xn[i];
                et[i] = (k < K1) ? xn[i] : ((k == K1) ? 0.0 : -INFINITY);
                mx = nanmax(mx, et[i]);
            }
            mx = warp_max(mx);
            // red[]: cnt, ssum, quad, data, (S d - a).p, ex.p — reduced together after the contraction
            double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                ex[i] = exp_noinline(et[i] - mx);
                if (k < K) { if (et[i] == mx) red[0] += 1.0; else red[1] += ex[i]; }
                if (k < TS) v0[k] = (k < K) ? ex[i] : 0.0;
                const double dk = xn[i] - muk[i];
                const double sdk = Sd[i] * dk;
                red[2] += sdk * dk;
                if (k < K1) { red[4] += (sdk - ak[i]) * pk[i]; red[5] += ex[i] * pk[i]; }
            }
            __syncwarp();
            SLT_T(t_e1);
            SLT_ACC(11, t_e0, t_e1);
            // data term: sum_v c_v (m + log(sum_k e_k beta_kv))           stm.py:938-941
            LogProd lp;
            logprod_init(lp);
            for (int w0 = 0; w0 < n; w0 += 32 * J) {
                double acc[J][2];
#pragma unroll
                for (int j = 0; j < J; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
                const double2* e2 = reinterpret_cast<const double2*>(v0);
                if (v.is_tm) {
                    // tile in TMEM: lane <-> word (slot j), 8 topics per tcgen05.ld (one pass: n <= 32 J)
                    int c0 = 0;
#pragma unroll 1
                    for (; c0 + 8 <= CS; c0 += 8) {
                        uint32_t b[J][8];
#pragma unroll
                        for (int j = 0; j < J; ++j) tm_ld8(v.taddr + (uint32_t)(j * CS + c0), b[j]);
                        const double2 e0 = e2[(c0 >> 1)], e1 = e2[(c0 >> 1) + 1], e2_ = e2[(c0 >> 1) + 2], e3 = e2[(c0 >> 1) + 3];
                        tm_wait_ld();
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            acc[j][0] = fma(e0.x, beta_u2d(b[j][0]), acc[j][0]);
                            acc[j][1] = fma(e0.y, beta_u2d(b[j][1]), acc[j][1]);
                            acc[j][0] = fma(e1.x, beta_u2d(b[j][2]), acc[j][0]);
                            acc[j][1] = fma(e1.y, beta_u2d(b[j][3]), acc[j][1]);
                            acc[j][0] = fma(e2_.x, beta_u2d(b[j][4]), acc[j][0]);
                            acc[j][1] = fma(e2_.y, beta_u2d(b[j][5]), acc[j][1]);
                            acc[j][0] = fma(e3.x, beta_u2d(b[j][6]), acc[j][0]);
                            acc[j][1] = fma(e3.y, beta_u2d(b[j][7]), acc[j][1]);
                        }
                    }
                    if ((CS - c0) & 4) {
                        uint32_t b[J][4];
#pragma unroll
                        for (int j = 0; j < J; ++j) tm_ld4(v.taddr + (uint32_t)(j * CS + c0), b[j]);
                        const double2 e0 = e2[(c0 >> 1)], e1 = e2[(c0 >> 1) + 1];
                        tm_wait_ld();
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            acc[j][0] = fma(e0.x, beta_u2d(b[j][0]), acc[j][0]);
                            acc[j][1] = fma(e0.y, beta_u2d(b[j][1]), acc[j][1]);
                            acc[j][0] = fma(e1.x, beta_u2d(b[j][2]), acc[j][0]);
                            acc[j][1] = fma(e1.y, beta_u2d(b[j][3]), acc[j][1]);
                        }
                        c0 += 4;
                    }
                    if ((CS - c0) & 2) {
                        uint32_t b[J][2];
#pragma unroll
                        for (int j = 0; j < J; ++j) tm_ld2(v.taddr + (uint32_t)(j * CS + c0), b[j]);
                        const double2 e0 = e2[(c0 >> 1)];
                        tm_wait_ld();
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            acc[j][0] = fma(e0.x, beta_u2d(b[j][0]), acc[j][0]);
                            acc[j][1] = fma(e0.y, beta_u2d(b[j][1]), acc[j][1]);
                        }
                    }
                } else {
                    const float4* rows[J];
#pragma unroll
                    for (int j = 0; j < J; ++j) {
                        int w = w0 + lane + 32 * j;
                        if (w >= n) w = n - 1;
                        rows[j] = reinterpret_cast<const float4*>(v.tile + (size_t)w * TS);
                    }
                    STM_UNROLL_Q
                    for (int q = 0; q < TS / 4; ++q) {
                        const double2 ea = e2[2 * q], eb = e2[2 * q + 1];
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const float4 b = rows[j][q];
                            acc[j][0] = fma(ea.x, beta_f2d(b.x), acc[j][0]);
                            acc[j][1] = fma(ea.y, beta_f2d(b.y), acc[j][1]);
                            acc[j][0] = fma(eb.x, beta_f2d(b.z), acc[j][0]);
                            acc[j][1] = fma(eb.y, beta_f2d(b.w), acc[j][1]);
                        }
                    }
                }
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const int w = w0 + lane + 32 * j;
                    if (w < n) logprod_add(lp, acc[j][0] + acc[j][1], v.cw[w]);
                }
                logprod_renorm(lp);
            }
            SLT_T(t_e2);
            SLT_ACC(12, t_e1, t_e2);
            red[3] = logprod_value(lp);
            warp_sum_n<6>(red);
            SLT_T(t_e3);
            SLT_ACC(13, t_e2, t_e3);
            if (lane == 0) {
                SlotBox* bx = v.box;
                bx->cur = nb; bx->hc[nb] = 1; bx->kind = 0;
                bx->res[0] = mx; bx->res[1] = red[0]; bx->res[2] = red[1]; bx->res[3] = red[2];
                bx->res[4] = red[3]; bx->res[5] = red[4]; bx->res[6] = red[5];
            }
        }
        __syncwarp();
        SLT_T(t_r2);
        SLT_ACC(1, t_r1, t_r2);

        // =================== scalar phase: lane g consumes slot g's evaluation =====================================
        int act = 0;    // 1: step accepted, 2: first evaluation of the document (initialise), 3: line search failed
        if (st == 1) {
            SlotBox* bx = view(lane).box;
            if (bx->kind == 0) {
                const double m = bx->res[0], cnt = bx->res[1];
                double ssum = bx->res[2];
                const double se_all = ssum + cnt;
                if (ssum != 0.0 && cnt != 1.0) ssum = ddiv(ssum, cnt);
                // scipy.special.logsumexp (scipy/special/_logsumexp.py:201-247)
                const double lse = log_noinline(1.0 + ssum) + ((cnt == 1.0) ? 0.0 : log_noinline(cnt)) + m;
                const double quad = 0.5 * bx->res[3];
                const double data = m * Nsum + bx->res[4];
                f_eval = quad - (data - Nint * lse);
                const double scale = ddiv(Nsum, se_all);
                dphi = bx->res[5] + scale * bx->res[6];   // = gt . p
                const int cur = bx->cur;
                bx->f[cur] = f_eval;
                bx->scale[cur] = scale;
                nfev++;
            } else {
                f_eval = bx->f[bx->cur];
                dphi = bx->res[7];
            }

            int accept = 0, fail = 0, start_w2 = 0, start_zoom = 0;
            double zl = 0, zh = 0, zpl = 0, zph = 0, zdl = 0;
            if (ls == LS_INIT) {
                act = 2;
            } else if (ls == LS_W1) {
                // one DCSRCH._iterate (scipy/optimize/_dcsrch.py:310-500) with (stp=alpha, f, g)
                const double stp_in = alpha, f = f_eval, gd = dphi;
                const double ftest = S.finit + stp_in * S.gtest;
                int warn = 0;
                if (S.stage == 1 && f <= ftest && gd >= 0.0) S.stage = 2;
                if (S.brackt && (stp_in <= S.stmin || stp_in >= S.stmax)) warn = 1;
                if (S.brackt && S.stmax - S.stmin <= xtol * S.stmax) warn = 1;
                if (stp_in == stpmax && f <= ftest && gd <= S.gtest) warn = 1;
                if (stp_in == stpmin && (f > ftest || gd >= S.gtest)) warn = 1;
                if (f <= ftest && fabs(gd) <= c2 * -S.ginit) {
                    accept = 1;
                } else if (warn) {
                    start_w2 = 1;
                } else {
                    double stp = stp_in;
                    {
                        // one call site on register copies: the modified function of stage 1
                        // (scipy/optimize/_dcsrch.py:448-470) only changes the operands
                        const bool mod = (S.stage == 1 && f <= S.fx && f > ftest);
                        const double gtest = S.gtest;
                        double stx = S.stx, sty = S.sty;
                        double fxm = S.fx, fym = S.fy, gxm = S.gx, gym = S.gy, fm = f, gm = gd;
                        int brackt = S.brackt;
                        if (mod) {
                            fm = f - stp * gtest; fxm = fxm - stx * gtest; fym = fym - sty * gtest;
                            gm = gd - gtest; gxm = gxm - gtest; gym = gym - gtest;
                        }
                        dcstep(stx, fxm, gxm, sty, fym, gym, stp, fm, gm, brackt, S.stmin, S.stmax);
                        if (mod) {
                            fxm = fxm + stx * gtest; fym = fym + sty * gtest;
                            gxm = gxm + gtest; gym = gym + gtest;
                        }
                        S.stx = stx; S.sty = sty; S.fx = fxm; S.fy = fym; S.gx = gxm; S.gy = gym; S.brackt = brackt;
                    }
                    if (S.brackt) {
                        if (fabs(S.sty - S.stx) >= 0.66 * S.width1) stp = S.stx + 0.5 * (S.sty - S.stx);
                        S.width1 = S.width;
                        S.width = fabs(S.sty - S.stx);
                        S.stmin = py_min2(S.stx, S.sty);
                        S.stmax = py_max2(S.stx, S.sty);
                    } else {
                        S.stmin = stp + 1.1 * (stp - S.stx);
                        S.stmax = stp + 4.0 * (stp - S.stx);
                    }
                    stp = np_clip(stp, stpmin, stpmax);
                    if ((S.brackt && (stp <= S.stmin || stp >= S.stmax)) ||
                        (S.brackt && S.stmax - S.stmin <= xtol * S.stmax))
                        stp = S.stx;
                    S.w1_it++;
                    if (!isfinite(stp) || S.w1_it >= 100) start_w2 = 1;  // WARN / maxiter -> stp None
                    else alpha = stp;
                }
            } else if (ls == LS_W2) {
                // bracket phase of scalar_search_wolfe2 (scipy/optimize/_linesearch.py:411-466)
                const double alpha1 = alpha, phi_a1 = f_eval, derphi_a1 = dphi;
                if (S.w2_i == 10) {
                    accept = 1;  // for-else: alpha_star = alpha1, derphi_star None (gradient re-evaluated)
                } else if (alpha1 == 0.0 || S.alpha0 > 1e100) {
                    fail = 1;
                } else if (phi_a1 > S.old_fval + c1 * alpha1 * S.derphi0 || (phi_a1 >= S.phi_a0 && S.w2_i > 0)) {
                    start_zoom = 1; zl = S.alpha0; zh = alpha1; zpl = S.phi_a0; zph = phi_a1; zdl = S.derphi_a0;
                } else if (fabs(derphi_a1) <= -c2 * S.derphi0) {
                    accept = 1;
                } else if (derphi_a1 >= 0.0) {
                    start_zoom = 1; zl = alpha1; zh = S.alpha0; zpl = phi_a1; zph = S.phi_a0; zdl = derphi_a1;
                } else {
                    const double alpha2 = py_min2(2.0 * alpha1, 1e100);
                    S.alpha0 = alpha1; S.phi_a0 = phi_a1; S.derphi_a0 = derphi_a1;
                    alpha = alpha2;
                    S.w2_i++;
                }
            } else {  // LS_ZOOM — scipy/optimize/_linesearch.py:546-634
                const double a_j = alpha, phi_aj = f_eval, derphi_aj = dphi;
                if (phi_aj > S.old_fval + c1 * a_j * S.derphi0 || phi_aj >= S.phi_lo) {
                    S.phi_rec = S.phi_hi; S.a_rec = S.a_hi; S.a_hi = a_j; S.phi_hi = phi_aj;
                } else {
                    if (fabs(derphi_aj) <= -c2 * S.derphi0) {
                        accept = 1;
                    } else {
                        if (derphi_aj * (S.a_hi - S.a_lo) >= 0.0) {
                            S.phi_rec = S.phi_hi; S.a_rec = S.a_hi; S.a_hi = S.a_lo; S.phi_hi = S.phi_lo;
                        } else {
                            S.phi_rec = S.phi_lo; S.a_rec = S.a_lo;
                        }
                        S.a_lo = a_j; S.phi_lo = phi_aj; S.derphi_lo = derphi_aj;
                    }
                }
                if (!accept) {
                    S.z_i++;
                    if (S.z_i > 10) fail = 1;
                }
            }
            if (start_w2) {
                // scalar_search_wolfe2 prologue (scipy/optimize/_linesearch.py:395-409)
                double alpha1;
                if (S.derphi0 != 0.0) alpha1 = py_min2(1.0, ddiv(1.01 * 2 * (S.old_fval - S.old_old_fval), S.derphi0));
                else alpha1 = 1.0;
                if (alpha1 < 0.0) alpha1 = 1.0;
                alpha1 = py_min2(alpha1, 1e100);
                S.alpha0 = 0.0; S.phi_a0 = S.old_fval; S.derphi_a0 = S.derphi0; S.w2_i = 0;
                alpha = alpha1;
                ls = LS_W2;
            }
            if (start_zoom) {
                S.a_lo = zl; S.a_hi = zh; S.phi_lo = zpl; S.phi_hi = zph; S.derphi_lo = zdl;
                S.phi_rec = S.old_fval; S.a_rec = 0.0; S.z_i = 0;
                ls = LS_ZOOM;
            }
            if (ls == LS_ZOOM && !accept && !fail) {
                // next trial step of _zoom
                const double dalpha = S.a_hi - S.a_lo;
                double za, zb;
                if (dalpha < 0.0) { za = S.a_hi; zb = S.a_lo; } else { za = S.a_lo; zb = S.a_hi; }
                double a_j = nan("");
                double cchk = 0.0;
                if (S.z_i > 0) {
                    cchk = 0.2 * dalpha;
                    a_j = cubicmin(S.a_lo, S.phi_lo, S.derphi_lo, S.a_hi, S.phi_hi, S.a_rec, S.phi_rec);
                }
                if (S.z_i == 0 || isnan(a_j) || a_j > zb - cchk || a_j < za + cchk) {
                    const double qchk = 0.1 * dalpha;
                    a_j = quadmin(S.a_lo, S.phi_lo, S.derphi_lo, S.a_hi, S.phi_hi);
                    if (isnan(a_j) || a_j > zb - qchk || a_j < za + qchk) a_j = S.a_lo + 0.5 * dalpha;
                }
                alpha = a_j;
            }
            if (fail) { warnflag = 2; act = 3; st = 2; }
            if (accept) {
                act = 1;
                k_it++;
                bx->k_it = k_it;
                bx->alpha = alpha;      // alpha_k of the accepted step (it is already there; kept explicit)
            }
        }
        __syncwarp();
        SLT_T(t_r3);
        SLT_ACC(2, t_r2, t_r3);

        // =================== accepted steps / first evaluations: BFGS update, new direction =========================
        // _minimize_bfgs body after the line search (scipy/optimize/_optimize.py:1452-1498)
        for (unsigned m = __ballot_sync(STM_FULL, act == 1 || act == 2); m != 0u; m &= m - 1u) {
            const int g = __ffs((int)m) - 1;
            const bool init = (__ballot_sync(STM_FULL, act == 2) >> g) & 1u;
            const SlotView v = view(g);
            const int cur = v.box->cur;
            const double sc = v.box->scale[cur];
            const double* xs = v.xt + cur * KS;
            const double* es = v.ex + cur * TS;
            double gt[KPL];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                gt[i] = 0.0;
                if (k < K1) {
                    const double dk = xs[k] - v.mu[k];
                    gt[i] = Sd[i] * dk - (v.a[k] - sc * es[k]);     // gradient stm.py:946-958 (reference quirk kept)
                }
            }
            if (init) {
                double n2 = 0.0, gm = 0.0, d_l = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = lane + 32 * i;
                    n2 += gt[i] * gt[i];
                    gm = nanmax(gm, fabs(gt[i]));
                    const double pi = -gt[i];                       // Hk = I
                    d_l += gt[i] * pi;
                    if (k < K1) { v.g[k] = gt[i]; v.p[k] = pi; }
                }
                n2 = warp_sum(n2);
                d_l = warp_sum(d_l);
                gm = warp_max(gm);
                if (lane == 0) { v.box->acc[0] = gm; v.box->acc[1] = d_l; v.box->acc[2] = n2; v.box->code = 0; }
                continue;
            }
            const double alpha_k = v.box->alpha;
            const int kit = v.box->k_it;
            double* Hk = P.scratch + (size_t)(gslot0 + g) * P.scratch_stride;   // BFGS inverse Hessian [K1][K1]
            double ys_l = 0.0, gm = 0.0, pm = 0.0;
            double sk[KPL], yk[KPL], pk[KPL];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                pk[i] = (k < K1) ? v.p[k] : 0.0;
                const double gold = (k < K1) ? v.g[k] : 0.0;
                sk[i] = __dmul_rn(alpha_k, pk[i]);
                if (k < K1) { v.x[k] = xs[k]; v.g[k] = gt[i]; }   // xk + alpha_k*pk, same expression as the trial point
                yk[i] = gt[i] - gold;
                ys_l += yk[i] * sk[i];
                gm = nanmax(gm, fabs(gt[i]));
                pm = nanmax(pm, fabs(pk[i]));
            }
            const double gnorm = warp_max(gm);
            pm = warp_max(pm);
            int code = 0;
            double derphi0 = 0.0;
            if (gnorm <= 1e-5) code = 1;
            else if (alpha_k * pm <= 0.0) code = 1;
            else if (!isfinite(v.box->f[cur])) code = 2;
            else {
                const double rhok_inv = warp_sum(ys_l);
                const double rho = (rhok_inv == 0.0) ? 1000.0 : ddiv(1.0, rhok_inv);
                // Hk <- (I - rho s y')(Hk)(I - rho y s') + rho s s'   as a symmetric rank-2 update:
                //   u = Hk y ;  Hk' = Hk - (rho u) s' - s (rho u)' + (rho^2 y'u + rho) s s'
                double u[KPL];
                double* t1 = v.xt + (cur ^ 1) * KS;      // the older memo point dies with the accepted step
                double* t2 = v.ex + (cur ^ 1) * TS;      // (only entries k < K-1 are touched: the zero pads stay)
                __syncwarp();
                if (kit == 1) {
#pragma unroll
                    for (int i = 0; i < KPL; ++i) u[i] = yk[i];  // Hk = I
                } else {
#pragma unroll
                    for (int i = 0; i < KPL; ++i) { const int k = lane + 32 * i; u[i] = 0.0; if (k < K1) t1[k] = yk[i]; }
                    __syncwarp();
                    for (int r = 0; r < K1; ++r) {
                        const double yr = t1[r];
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < K1) u[i] = fma(Hk[(size_t)r * K1 + k], yr, u[i]);
                        }
                    }
                    __syncwarp();
                }
                double yu_l = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) yu_l += yk[i] * u[i];
                const double yu = warp_sum(yu_l);
                const double cc = rho * rho * yu + rho;
                double ru[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = lane + 32 * i;
                    ru[i] = rho * u[i];
                    if (k < K1) { t1[k] = sk[i]; t2[k] = ru[i]; }
                }
                __syncwarp();
                // fused: write Hk' and accumulate p = -Hk' g
                double pn[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) pn[i] = 0.0;
                for (int r = 0; r < K1; ++r) {
                    const double sr = t1[r], rur = t2[r], gr = v.g[r];
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int k = lane + 32 * i;
                        if (k < K1) {
                            double h = (kit == 1) ? ((r == k) ? 1.0 : 0.0) : Hk[(size_t)r * K1 + k];
                            h = h - (__dmul_rn(rur, sk[i]) + __dmul_rn(sr, ru[i])) + cc * sr * sk[i];
                            Hk[(size_t)r * K1 + k] = h;
                            pn[i] = fma(h, gr, pn[i]);
                        }
                    }
                }
                __syncwarp();
                double d_l = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = lane + 32 * i;
                    const double pi = -pn[i];
                    if (k < K1) v.p[k] = pi;
                    d_l += gt[i] * pi;
                }
                derphi0 = warp_sum(d_l);
                // the older memo buffer was used as scratch: invalidate it
                if (lane == 0) v.box->hc[cur ^ 1] = 0;
            }
            if (lane == 0) { v.box->acc[0] = gnorm; v.box->acc[1] = derphi0; v.box->code = code; }
        }
        __syncwarp();
        SLT_T(t_r4);
        SLT_ACC(3, t_r3, t_r4);

        // =================== scalar phase 2: start the next line search / finish =================================
        if (act == 1 || act == 2) {
            SlotBox* bx = view(lane).box;
            int done = 0;
            if (act == 2) {
                S.old_fval = f_eval;
                S.old_old_fval = S.old_fval + dsqrt(bx->acc[2]) * 0.5;
                S.gnorm = bx->acc[0];
            } else {
                S.old_old_fval = S.old_fval;
                S.old_fval = f_eval;
                S.gnorm = bx->acc[0];
                const int code = bx->code;
                if (code == 1) done = 1;
                else if (code == 2) { warnflag = 2; done = 1; }
            }
            if (!done) {
                if (!(S.gnorm > 1e-5) || !(k_it < maxiter)) {
                    done = 1;
                } else {
                    S.derphi0 = bx->acc[1];
                    // scalar_search_wolfe1 prologue + DCSRCH START
                    double alpha1;
                    if (S.derphi0 != 0.0) {
                        alpha1 = py_min2(1.0, ddiv(1.01 * 2 * (S.old_fval - S.old_old_fval), S.derphi0));
                        if (alpha1 < 0.0) alpha1 = 1.0;
                    } else alpha1 = 1.0;
                    if (alpha1 < stpmin || alpha1 > stpmax || S.derphi0 >= 0.0 || !isfinite(alpha1)) {
                        // task = ERROR -> stp None -> wolfe2
                        double a1 = py_min2(alpha1, 1e100);
                        S.alpha0 = 0.0; S.phi_a0 = S.old_fval; S.derphi_a0 = S.derphi0; S.w2_i = 0;
                        alpha = a1;
                        ls = LS_W2;
                    } else {
                        S.brackt = 0; S.stage = 1; S.finit = S.old_fval; S.ginit = S.derphi0; S.gtest = c1 * S.ginit;
                        S.width = stpmax - stpmin; S.width1 = S.width * 2.0;
                        S.stx = 0.0; S.fx = S.finit; S.gx = S.ginit; S.sty = 0.0; S.fy = S.finit; S.gy = S.ginit;
                        S.stmin = 0.0; S.stmax = alpha1 + 4.0 * alpha1;
                        S.w1_it = 1;  // the START call was iteration 0 of DCSRCH.__call__
                        alpha = alpha1;
                        ls = LS_W1;
                    }
                }
            }
            if (done) st = 2;
        }
        if (st == 1) view(lane).box->alpha = alpha;
        __syncwarp();
        SLT_T(t_r5);
        SLT_ACC(4, t_r4, t_r5);

        // =================== finished documents: eta, status =====================================================
        for (unsigned m = __ballot_sync(STM_FULL, st == 2); m != 0u; m &= m - 1u) {
            const int g = __ffs((int)m) - 1;
            const SlotView v = view(g);
            const int d = v.box->d;
            bool bad = false;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                if (k < K1) {
                    const double xv = v.x[k];
                    P.eta[(size_t)d * K1 + k] = xv;
                    bad = bad || isnan(xv);
                }
            }
            const bool badx = __any_sync(STM_FULL, bad);
            if (lane == g) {
                int status = warnflag;
                if (status != 2) {
                    if (k_it >= maxiter) status = 1;
                    else status = (badx || isnan(S.gnorm) || isnan(S.old_fval)) ? 3 : 0;
                }
                const int nit_c = k_it > 0xfffff ? 0xfffff : k_it;
                P.doc_info[d] = status | (nit_c << 4);
                P.doc_nfev[d] = nfev;
                st = 0;
            }
        }
        __syncwarp();
        SLT_T(t_r6);
        SLT_ACC(5, t_r5, t_r6);
    }
#if STM_SLOTS_TIMING
    if (lane == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(P.dbg_cycles + i, (unsigned long long)dbg_t[i]);
#endif
    if (TW > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_s));
    }
}

}  // namespace stm
