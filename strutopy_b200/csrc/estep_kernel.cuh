// estep_kernel.cuh — the STM E-step as ONE sm_100a kernel: per-document Laplace-variational inference.
//
// Replaces the document loop of the reference, /root/reference/src/modules/stm.py:519-590
// (get_beta :599, optimize_eta :917 -> scipy BFGS, hessian :986, make_pd :964, decompose_hessian
// :1031, lower_bound :1068, optimize_nu :1052, update_z :1103, accumulation :582-590).
//
// Mapping:
//   * one WARP per document, persistent warps pulling document indices from an atomic queue;
//     control flow (SciPy's BFGS / dcsrch / zoom state machine) is warp-uniform, so different
//     documents diverge across warps for free;
//   * the document's beta rows (fp32, word-major [V][TS]) are gathered ONCE into shared memory by
//     TMA bulk copies (cp.async.bulk, one 16B-aligned row per word, completion on an mbarrier);
//   * every objective evaluation is a (K x n_d) contraction out of shared memory with fp64
//     accumulation, warp-shuffle reductions, fp64 line-search scalars;
//   * K-vectors (eta, p, g, ...) live in registers, lane l owning k = l, l+32, ...;
//   * the dense BFGS inverse-Hessian lives in a per-warp L2-resident scratch (touched once per
//     accepted step); the Laplace Hessian is accumulated in register blocks, then factorised in the
//     shared memory the beta tile occupied;
//   * phi is scattered into beta_ss with fp64 reductions (red.global.add.f64), nu into replicated
//     sigma_ss accumulators.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <math.h>

// ---- tuning switches (A/B builds: make variant NAME=.. DEFS=..) ---------------------------------
#ifndef STM_LOGPROD
#define STM_LOGPROD 1      // 1: sum_v c_v log s_v as one log of a product per lane; 0: one log per word
#endif
#ifndef STM_Q_UNROLL
#define STM_Q_UNROLL 1     // unroll factor of the contraction's k-group loop; 0 = compiler default
#endif
#ifndef STM_INT_CVT
#define STM_INT_CVT 1      // 1: integer fp32->fp64 conversion of beta (IMAD.WIDE); 0: F2F (XU pipe)
#endif
#ifndef STM_LOG_NOINLINE
#define STM_LOG_NOINLINE 1 // keep log()/exp() slow paths out of line (I-cache footprint)
#endif
// development-only switches for time attribution (results are WRONG when any is set)
#ifndef STM_DCSTEP_INLINE
#define STM_DCSTEP_INLINE 1   // 1: dcstep inlined at ONE call site on register copies of the dcsrch state (r01 A/B: kernel A 33.4 -> 31.2 ms)
#endif
#ifndef STM_UNIFORM_LS
#define STM_UNIFORM_LS 1      // 1: make the line-search inputs PROVABLY warp-uniform (redux.sync), so that the scalar
#endif                        //    state machine compiles to uniform branches without convergence barriers
#ifndef STM_LS_REGS
#define STM_LS_REGS 0         // 1: line-search scalars in registers instead of shared memory (needs a larger register budget)
#endif
#ifndef STM_DDIV_INLINE
#define STM_DDIV_INLINE 0     // 1: fp64 divisions of the line-search code inlined (ILP between independent quotients)
#endif
#ifndef STM_DBG_NO_PHI
#define STM_DBG_NO_PHI 0
#endif
#ifndef STM_DBG_SKIP_POST
#define STM_DBG_SKIP_POST 0
#endif
#ifndef STM_DBG_SKIP_BFGS
#define STM_DBG_SKIP_BFGS 0
#endif
#ifndef STM_DBG_TIMING
#define STM_DBG_TIMING 0   // 1: accumulate clock64() per phase into P.dbg_cycles[8] (variant builds only)
#endif
#ifndef STM_DBG_SKIP_DENSE
#define STM_DBG_SKIP_DENSE 0
#endif
#ifndef STM_BFGS_MAX_THREADS
#ifndef STM_CURV_CERT
#define STM_CURV_CERT 1       // 1: end a line search whose bracket lies where the curvature condition provably cannot hold (see
#endif                        //    the comment at "curvature certificate" in kernel A)
#define STM_BFGS_MAX_THREADS 384   // launch bound of kernel A (register budget = 65536 / this): 12 warps x 160 registers, no spills (r01 A/B: 448 -> 384: 31.3 -> 29.3 ms)
#endif
#define STM_PRAGMA2_(x) _Pragma(#x)
#define STM_PRAGMA_(x) STM_PRAGMA2_(x)
#if STM_Q_UNROLL > 0
#define STM_UNROLL_Q STM_PRAGMA_(unroll STM_Q_UNROLL)
#else
#define STM_UNROLL_Q
#endif
#if STM_LOG_NOINLINE
#define STM_NOINLINE __noinline__
#else
#define STM_NOINLINE __forceinline__
#endif

namespace stm {

struct EstepParams {
    // corpus (CSR), resident in HBM for the whole fit
    const long long* doc_ptr;   // [D+1]
    const int* word_id;         // [nnz]
    const float* count;         // [nnz]
    const int* aspect;          // [D] or nullptr
    // work list of this launch (documents of one length class)
    const int* docs;            // [n_docs] document indices
    int n_docs;
    unsigned int* queue;        // atomic work counter (zeroed before launch)
    // model
    int K, V, A;
    int TS;                     // beta row stride in floats (multiple of 4, TS/4 odd)
    const float* beta_t;        // [A][V][TS] word-major beta
    const double* mu;           // [D][K-1]
    const double* prior;        // [K-1] diagonal of siginv (stm.py:501 makes it diagonal), then sigmaentropy
    // per-document state / outputs
    double* eta;                // [D][K-1] in: warm start, out: result
    double* theta;              // [D][K]
    double* doc_bound;          // [D]
    int* doc_info;              // [D] status | nit<<4 | repair<<24   (nit clipped to 20 bits)
    int* doc_nfev;              // [D] objective evaluations actually performed
    // sufficient statistics
    double* beta_ss_t;          // [A][V][TS] fp64, accumulated
    double* sigma_ss_rep;       // [n_rep][(K-1)*(K-1)] replicated accumulators (lower triangle used)
    int n_rep;
    // per-warp scratch in HBM/L2: 2 * HS*K1 doubles (BFGS inverse Hessian, Laplace Hessian)
    double* scratch;
    long long scratch_stride;   // doubles per warp
    // shared memory geometry
    int n_cap;                  // tile rows (words) per warp
    int smem_per_warp;          // bytes of the per-warp tile block (kernel A: smem warps only)
    int smem_small;             // kernel A: bytes of the per-warp small block (K-vectors, line-search state)
    int tm_warps;               // kernel A: warps 0..tm_warps-1 keep their tile in TMEM (0, 4 or 8)
    int tm_cols;                // kernel A: TMEM columns per TMEM warp (512 or 256)
    unsigned long long* dbg_cycles;  // [8] phase cycle counters (only written when STM_DBG_TIMING)
};
#if STM_DBG_TIMING
#define STM_T(var) const long long var = clock64()
#define STM_TACC(slot, t0, t1) dbg_t[slot] += (t1) - (t0)
#else
#define STM_T(var)
#define STM_TACC(slot, t0, t1)
#endif

#define STM_FULL 0xffffffffu

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(STM_FULL, v, o);
    return v;
}
// N independent sums in one butterfly: the shuffles of different values pipeline, so the latency is
// that of ONE reduction
template <int N>
__device__ __forceinline__ void warp_sum_n(double (&v)[N]) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        double t[N];
#pragma unroll
        for (int i = 0; i < N; ++i) t[i] = __shfl_xor_sync(STM_FULL, v[i], o);
#pragma unroll
        for (int i = 0; i < N; ++i) v[i] += t[i];
    }
}
__device__ __forceinline__ double nanmax(double a, double b) {
    return (isnan(a) || isnan(b)) ? nan("") : fmax(a, b);
}
__device__ __forceinline__ double warp_max(double v) {
    const bool has_nan = __any_sync(STM_FULL, isnan(v));  // np.max propagates NaN
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(STM_FULL, v, o));
    return has_nan ? nan("") : v;
}

// Python / NumPy comparison semantics used by SciPy's line searches (see oracle/stm_oracle.c)
__device__ __forceinline__ double py_max3(double a, double b, double c) {
    double m = a;
    if (b > m) m = b;
    if (c > m) m = c;
    return m;
}
__device__ __forceinline__ double py_max2(double a, double b) { return (b > a) ? b : a; }
__device__ __forceinline__ double py_min2(double a, double b) { return (b < a) ? b : a; }
__device__ __forceinline__ double np_clip(double x, double lo, double hi) {
    if (isnan(x)) return x;
    double t = (x < lo) ? lo : x;
    return (t > hi) ? hi : t;
}
__device__ __forceinline__ double np_sign(double x) {
    if (isnan(x)) return x;
    return (x > 0.0) ? 1.0 : ((x < 0.0) ? -1.0 : 0.0);
}

// A value that is identical in every lane, passed through redux.sync so that the compiler's divergence analysis
// knows it (AND of equal bit patterns is the identity).  Branches on such values need no BSSY/BSYNC scaffolding.
__device__ __forceinline__ double uniform_f64(double v) {
    const int hi = (int)__reduce_and_sync(STM_FULL, (unsigned)__double2hiint(v));
    const int lo = (int)__reduce_and_sync(STM_FULL, (unsigned)__double2loint(v));
    return __hiloint2double(hi, lo);
}

// IEEE fp64 division / square root kept out of line in the (scalar, warp-uniform) line-search code:
// each inlined copy is ~25 instructions and the state machine has dozens of them (I-cache footprint).
#if STM_DDIV_INLINE
static __device__ __forceinline__ double ddiv(double a, double b) { return a / b; }
#else
static __device__ __noinline__ double ddiv(double a, double b) { return a / b; }
#endif
static __device__ __noinline__ double dsqrt(double a) { return sqrt(a); }

// MINPACK-2 dcstep — scipy/optimize/_dcsrch.py:502-728.  The four cases share ONE copy of the cubic
// interpolation arithmetic (theta, s, gamma, r): each case performs exactly the operations of the
// SciPy source in the same order, only the operands are selected first (instruction footprint).
#if STM_DCSTEP_INLINE
static __device__ __forceinline__ void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy,
#else
static __device__ __noinline__ void dcstep(double& stx, double& fx, double& dx, double& sty, double& fy,
#endif
                                    double& dy, double& stp, double fp, double dp, int& brackt,
                                    double stpmin, double stpmax) {
    const double sgnd = np_sign(dp) * np_sign(dx);
    int cs;  // 1..4: the four Moré–Thuente cases; 0: case 4 without a bracket (no interpolation)
    if (fp > fx) cs = 1;
    else if (sgnd < 0.0) cs = 2;
    else if (fabs(dp) < fabs(dx)) cs = 3;
    else cs = brackt ? 4 : 0;
    double stpf;
    if (cs != 0) {
        // theta = 3 (f1 - f2) / (s2 - s1) + d1 + dp
        const double f1 = (cs == 4) ? fp : fx, f2 = (cs == 4) ? fy : fp;
        const double s2 = (cs == 4) ? sty : stp, s1 = (cs == 4) ? stp : stx;
        const double d1 = (cs == 4) ? dy : dx;
        const double theta = ddiv(3.0 * (f1 - f2), s2 - s1) + d1 + dp;
        const double s = py_max3(fabs(theta), fabs(d1), fabs(dp));
        const double ts = ddiv(theta, s);
        double arg = ts * ts - ddiv(d1, s) * ddiv(dp, s);
        if (cs == 3) arg = py_max2(0.0, arg);
        double gamma = s * dsqrt(arg);
        const bool flip = (cs == 1) ? (stp < stx) : ((cs == 4) ? (stp > sty) : (stp > stx));
        if (flip) gamma = -gamma;
        double p, q;
        if (cs == 1) { p = (gamma - dx) + theta; q = ((gamma - dx) + gamma) + dp; }
        else if (cs == 2) { p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dx; }
        else if (cs == 3) { p = (gamma - dp) + theta; q = (gamma + (dx - dp)) + gamma; }
        else { p = (gamma - dp) + theta; q = ((gamma - dp) + gamma) + dy; }
        const double r = ddiv(p, q);
        if (cs == 1) {
            const double stpc = stx + r * (stp - stx);
            const double stpq = stx + (ddiv(dx, ddiv(fx - fp, stp - stx) + dx) * 0.5) * (stp - stx);
            if (fabs(stpc - stx) <= fabs(stpq - stx)) stpf = stpc;
            else stpf = stpc + (stpq - stpc) * 0.5;
            brackt = 1;
        } else if (cs == 4) {
            stpf = stp + r * (sty - stp);
        } else {
            const double stpq = stp + ddiv(dp, dp - dx) * (stx - stp);
            if (cs == 2) {
                const double stpc = stp + r * (stx - stp);
                if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
                else stpf = stpq;
                brackt = 1;
            } else {
                double stpc;
                if (r < 0.0 && gamma != 0.0) stpc = stp + r * (stx - stp);
                else if (stp > stx) stpc = stpmax;
                else stpc = stpmin;
                if (brackt) {
                    if (fabs(stpc - stp) < fabs(stpq - stp)) stpf = stpc;
                    else stpf = stpq;
                    if (stp > stx) stpf = py_min2(stp + 0.66 * (sty - stp), stpf);
                    else stpf = py_max2(stp + 0.66 * (sty - stp), stpf);
                } else {
                    if (fabs(stpc - stp) > fabs(stpq - stp)) stpf = stpc;
                    else stpf = stpq;
                    stpf = np_clip(stpf, stpmin, stpmax);
                }
            }
        }
    } else if (stp > stx) stpf = stpmax;
    else stpf = stpmin;

    if (fp > fx) {
        sty = stp; fy = fp; dy = dp;
    } else {
        if (sgnd < 0.0) { sty = stx; fy = fx; dy = dx; }
        stx = stp; fx = fp; dx = dp;
    }
    stp = stpf;
}

// scipy/optimize/_linesearch.py:491-543; NaN stands for None
static __device__ __noinline__ double cubicmin(double a, double fa, double fpa, double b, double fb,
                                        double c, double fc) {
    const double C = fpa, db = b - a, dc = c - a;
    const double denom = (db * dc) * (db * dc) * (db - dc);
    const double d00 = dc * dc, d01 = -(db * db), d10 = -(dc * dc * dc), d11 = db * db * db;
    const double v0 = fb - fa - C * db, v1 = fc - fa - C * dc;
    double A = __dadd_rn(__dmul_rn(d00, v0), __dmul_rn(d01, v1));
    double B = __dadd_rn(__dmul_rn(d10, v0), __dmul_rn(d11, v1));
    if (denom == 0.0) return nan("");
    A = ddiv(A, denom);
    B = ddiv(B, denom);
    const double radical = __dadd_rn(__dmul_rn(B, B), -__dmul_rn(3.0 * A, C));
    if (radical < 0.0 || A == 0.0) return nan("");
    const double xmin = a + ddiv(-B + dsqrt(radical), 3.0 * A);
    if (!isfinite(xmin)) return nan("");
    return xmin;
}
static __device__ __noinline__ double quadmin(double a, double fa, double fpa, double b, double fb) {
    const double D = fa, C = fpa, db = b - a * 1.0;
    if (db * db == 0.0) return nan("");
    const double B = ddiv(__dadd_rn(__dadd_rn(fb, -D), -__dmul_rn(C, db)), db * db);
    if (2.0 * B == 0.0) return nan("");
    const double xmin = a - ddiv(C, 2.0 * B);
    if (!isfinite(xmin)) return nan("");
    return xmin;
}

// fp32 -> fp64 of a NON-NEGATIVE beta entry with ONE integer multiply-add (IMAD.WIDE) instead of
// F2F.F64.F32: conversions run on the XU pipe at 1/8 rate and saturated it (profiles/r01b).
// bits64 = bits32 * 2^29 + (896 << 52): exact for every normal float; 0 and subnormals
// (< 1.18e-38) map to [2^-127, 2^-126), i.e. an absolute error below 1.2e-38.
__device__ __forceinline__ double beta_f2d(float f) {
#if STM_INT_CVT
    const unsigned long long r =
        (unsigned long long)__float_as_uint(f) * 0x20000000ull + 0x3800000000000000ull;
    return __longlong_as_double((long long)r);
#else
    return (double)f;
#endif
}

// runtime-indexed read of a small register array (compiles to a select chain; multi-pass DMMA only)
template <int N>
__device__ __forceinline__ double fr_sel(const double (&fr)[N], int idx) {
    double r = fr[0];
#pragma unroll
    for (int t = 1; t < N; ++t) r = (idx == t) ? fr[t] : r;
    return r;
}

// log() kept out of line: it is only needed once per lane per evaluation (see logprod_*), so one
// shared copy keeps the hot loop's instruction footprint small (the kernel is I-cache sensitive).
static __device__ STM_NOINLINE double log_noinline(double x) { return log(x); }
static __device__ STM_NOINLINE double exp_noinline(double x) { return exp(x); }

// sum_v c_v log(s_v) accumulated as a PRODUCT: s = m2 * 2^e (m2 in [1,2)), prod *= m2^c, esum += c*e.
// One log per lane per evaluation instead of one per word.  The inlined fast path handles c = 1 and a
// positive normal s (almost every word); everything else (c = 2..8, non-integer / large counts, zero,
// subnormal, negative or non-finite s) goes through ONE out-of-line copy (instruction footprint).
struct LogProd {
    double prod;   // running mantissa product, kept in [1, 2^64)
    int esum;      // running exponent sum
    double extra;  // fallback terms c*log(s)
};
struct LogProdTerm { double pw; double extra; int e; };
__device__ __forceinline__ void logprod_init(LogProd& a) { a.prod = 1.0; a.esum = 0; a.extra = 0.0; }
__device__ __forceinline__ void logprod_renorm(LogProd& a) {
    const int hi = __double2hiint(a.prod);
    a.esum += ((hi >> 20) & 0x7ff) - 1023;
    a.prod = __hiloint2double((hi & 0x800fffff) | 0x3ff00000, __double2loint(a.prod));
}
static __device__ __noinline__ LogProdTerm logprod_slow(double s, float cf) {
    LogProdTerm t;
    const int hi = __double2hiint(s);
    const int be = (hi >> 20) & 0x7ff;           // biased exponent; sign bit excluded below
    const int ci = (int)cf;
    const bool fast = (hi > 0) && (be != 0) && (be != 0x7ff) && ((float)ci == cf) && (ci >= 1) && (ci <= 8);
    if (fast) {
        const double m2 = __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
        // m2^ci for ci in 1..8 without a loop: m2^(b0) * (m2^2)^(b1) * (m2^4)^(b2) * (m2^8)^(b3)
        const double q2 = m2 * m2, q4 = q2 * q2;
        double pw = (ci & 1) ? m2 : 1.0;
        if (ci & 2) pw *= q2;
        if (ci & 4) pw *= q4;
        if (ci & 8) pw *= q4 * q4;
        t.pw = pw; t.e = ci * (be - 1023); t.extra = 0.0;
    } else {
        t.pw = 1.0; t.e = 0; t.extra = (double)cf * log_noinline(s);
    }
    return t;
}
__device__ __forceinline__ void logprod_add(LogProd& a, double s, float cf) {
    const int hi = __double2hiint(s);
    const unsigned se = (unsigned)hi >> 20;      // sign | biased exponent: positive normal iff 1 <= se <= 0x7fe
    if (cf == 1.0f && (se - 1u) < 0x7feu) {
        a.prod *= __hiloint2double((hi & 0x000fffff) | 0x3ff00000, __double2loint(s));
        a.esum += (int)se - 1023;
    } else {
        const LogProdTerm t = logprod_slow(s, cf);
        a.prod *= t.pw; a.esum += t.e; a.extra += t.extra;
    }
}
__device__ __forceinline__ double logprod_value(const LogProd& a) {
    return (double)a.esum * 0.6931471805599453 + log_noinline(a.prod) + a.extra;
}

// ---- TMA bulk copy + mbarrier (PTX) -----------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, int count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    asm volatile(
        "{\n"
        ".reg .pred P1;\n"
        "LAB_WAIT:\n"
        "mbarrier.try_wait.parity.shared::cta.b64 P1, [%0], %1;\n"
        "@P1 bra DONE;\n"
        "bra LAB_WAIT;\n"
        "DONE:\n"
        "}\n" ::"r"(smem_u32(bar)), "r"(parity)
        : "memory");
}
__device__ __forceinline__ void tma_row_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst)),
        "l"(src), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void fence_proxy_async() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void red_add_f64(double* addr, double v) {
    asm volatile("red.global.add.f64 [%0], %1;" ::"l"(addr), "d"(v) : "memory");
}

// ---- tensor memory (TMEM) as tile storage: tcgen05.st / tcgen05.ld, shape 32x32b (thread l of the
// warp <-> TMEM lane 32*(warp%4)+l, registers <-> consecutive columns) ----------------------------
__device__ __forceinline__ void tm_st8(uint32_t taddr, const float4& a, const float4& b) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
                 "r"(__float_as_uint(a.w)), "r"(__float_as_uint(b.x)), "r"(__float_as_uint(b.y)),
                 "r"(__float_as_uint(b.z)), "r"(__float_as_uint(b.w))
                 : "memory");
}
__device__ __forceinline__ void tm_st4(uint32_t taddr, const float4& a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1,%2,%3,%4};" ::"r"(taddr),
                 "r"(__float_as_uint(a.x)), "r"(__float_as_uint(a.y)), "r"(__float_as_uint(a.z)),
                 "r"(__float_as_uint(a.w))
                 : "memory");
}
__device__ __forceinline__ void tm_st2(uint32_t taddr, const float2& a) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x2.b32 [%0], {%1,%2};" ::"r"(taddr), "r"(__float_as_uint(a.x)),
                 "r"(__float_as_uint(a.y))
                 : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_ld4(uint32_t taddr, uint32_t (&v)[4]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_ld2(uint32_t taddr, uint32_t (&v)[2]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x2.b32 {%0,%1}, [%2];" : "=r"(v[0]), "=r"(v[1]) : "r"(taddr) : "memory");
}
// four doubles of this lane <-> 8 consecutive 32-bit columns of its tensor-memory lane
__device__ __forceinline__ void tm_st_d4(uint32_t taddr, double a, double b, double c, double d) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr),
                 "r"(__double2loint(a)), "r"(__double2hiint(a)), "r"(__double2loint(b)), "r"(__double2hiint(b)),
                 "r"(__double2loint(c)), "r"(__double2hiint(c)), "r"(__double2loint(d)), "r"(__double2hiint(d))
                 : "memory");
}
__device__ __forceinline__ void tm_d4_from(const uint32_t (&v)[8], double& a, double& b, double& c, double& d) {
    a = __hiloint2double((int)v[1], (int)v[0]); b = __hiloint2double((int)v[3], (int)v[2]);
    c = __hiloint2double((int)v[5], (int)v[4]); d = __hiloint2double((int)v[7], (int)v[6]);
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
// beta_f2d on raw bits (what tcgen05.ld returns)
__device__ __forceinline__ double beta_u2d(uint32_t u) { return beta_f2d(__uint_as_float(u)); }

// ---- dense (K-1)x(K-1) helpers on a shared-memory matrix, one warp, out of line -------------------
// Storage convention: Hm is [K1][HS] (HS odd).  The UPPER triangle keeps the original symmetric
// matrix; the strict LOWER triangle receives the Cholesky factor; its diagonal goes to Ld.

// left-looking Cholesky, lane <-> row, inner products as 4 interleaved partial sums; the pivot
// H_jj - sum_k<j L_jk^2 is kept as a running diagonal Dw (same subtraction order as a sequential loop).
static __device__ __noinline__ int chol_factor(double* Hm, double* Dw, const double* Dg, double* Ld,
                                               int K1, int HS, int lane) {
    for (int k = lane; k < K1; k += 32) Dw[k] = Dg[k];
    __syncwarp();
    for (int j = 0; j < K1; ++j) {
        const double djj = Dw[j];
        if (!(djj > 0.0)) return 0;
        const double ljj = sqrt(djj);
        const double inv = 1.0 / ljj;
        if (lane == 0) Ld[j] = ljj;
        const double* Lj = Hm + (size_t)j * HS;
        for (int i = j + 1 + lane; i < K1; i += 32) {
            const double* Li_ = Hm + (size_t)i * HS;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int k = 0;
            for (; k + 4 <= j; k += 4) {
                t0 = fma(Li_[k], Lj[k], t0);
                t1 = fma(Li_[k + 1], Lj[k + 1], t1);
                t2 = fma(Li_[k + 2], Lj[k + 2], t2);
                t3 = fma(Li_[k + 3], Lj[k + 3], t3);
            }
            for (; k < j; ++k) t0 = fma(Li_[k], Lj[k], t0);
            const double lij = (Lj[i] - ((t0 + t1) + (t2 + t3))) * inv;  // upper triangle holds H_ij
            Hm[(size_t)i * HS + j] = lij;
            Dw[i] = Dw[i] - lij * lij;
        }
        __syncwarp();
    }
    return 1;
}

// nu = H^-1 = L^-T L^-1 (stm.py:1052-1066), accumulated into sigma_ss (stm.py:582).
// Li = L^-1 (lower) is stored transposed in the UPPER triangle: Hm[j][i] = Li[i][j], i >= j.
//   row i for all columns j < i in parallel (lane <-> j):
//   Li[i][j] = -(sum_{k=j}^{i-1} L[i][k] Li[k][j]) / L[i][i],   Li[i][i] = 1 / L[i][i]
static __device__ __noinline__ void inverse_and_nu(double* Hm, const double* Ld, double* sig_acc, int K1,
                                                   int HS, int lane) {
    for (int i = 0; i < K1; ++i) {
        const double inv = 1.0 / Ld[i];
        const double* Li_ = Hm + (size_t)i * HS;
        if (lane == 0) Hm[(size_t)i * HS + i] = inv;
        for (int j = lane; j < i; j += 32) {
            const double* Uj = Hm + (size_t)j * HS;
            double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
            int k = j;
            for (; k + 4 <= i; k += 4) {
                t0 = fma(Li_[k], Uj[k], t0);
                t1 = fma(Li_[k + 1], Uj[k + 1], t1);
                t2 = fma(Li_[k + 2], Uj[k + 2], t2);
                t3 = fma(Li_[k + 3], Uj[k + 3], t3);
            }
            for (; k < i; ++k) t0 = fma(Li_[k], Uj[k], t0);
            Hm[(size_t)j * HS + i] = -((t0 + t1) + (t2 + t3)) * inv;
        }
        __syncwarp();
    }
    // nu_ij = sum_{k>=i} Li[k][i] Li[k][j]  (i >= j): row i of the upper storage dotted with row j
    int i = 0;
    for (int idx = lane; idx < K1 * (K1 + 1) / 2; idx += 32) {
        while ((i + 1) * (i + 2) / 2 <= idx) i++;
        const int j = idx - i * (i + 1) / 2;
        const double* Ui = Hm + (size_t)i * HS;
        const double* Uj = Hm + (size_t)j * HS;
        double t0 = 0.0, t1 = 0.0, t2 = 0.0, t3 = 0.0;
        int k = i;
        for (; k + 4 <= K1; k += 4) {
            t0 = fma(Ui[k], Uj[k], t0);
            t1 = fma(Ui[k + 1], Uj[k + 1], t1);
            t2 = fma(Ui[k + 2], Uj[k + 2], t2);
            t3 = fma(Ui[k + 3], Uj[k + 3], t3);
        }
        for (; k < K1; ++k) t0 = fma(Ui[k], Uj[k], t0);
        red_add_f64(sig_acc + (size_t)i * K1 + j, (t0 + t1) + (t2 + t3));
    }
}

// status / repair codes mirror oracle/stm_oracle.h
enum { LS_INIT = 0, LS_W1 = 1, LS_W2 = 2, LS_ZOOM = 3 };

// ===== BFGS KERNEL =====
// Line-search / BFGS scalars (warp-uniform) live in shared memory, not registers: the state machine
// touches them only between evaluations, and the registers they would pin (~60) are what limits
// the number of resident documents per SM.  Every lane writes the same value; reads are broadcasts.
struct LsState {
    double old_fval, old_old_fval, gnorm, derphi0, f2;
    // dcsrch (scipy/optimize/_dcsrch.py)
    double finit, ginit, gtest, width, width1, stx, fx, gx, sty, fy, gy, stmin, stmax;
    // scalar_search_wolfe2 / _zoom (scipy/optimize/_linesearch.py)
    double alpha0, phi_a0, derphi_a0, a_lo, a_hi, phi_lo, phi_hi, derphi_lo, phi_rec, a_rec;
    double a_safe;   // curvature certificate of the current direction (STM_CURV_CERT), 0 = none
    int brackt, stage, w1_it, w2_i, z_i, pad_;
};

// Kernel A: per-document BFGS (stm.py:536-545 -> scipy.optimize.minimize(method="BFGS")).
// Writes eta (in place), doc_info = status | nit << 4, doc_nfev.
template <int KPL, int J>
// Registers are allocated in groups of 4 warps: 9-12 warps get 168 registers per thread, 13-16 warps only 128
// (which spills), so 12 warps is the widest configuration without spills (r01 A/B in profiles/r01_tuning_log.md).
__global__ void __launch_bounds__(STM_BFGS_MAX_THREADS, 1) bfgs_kernel(const EstepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
#if STM_UNIFORM_LS
    const int warp = __reduce_max_sync(STM_FULL, (int)(threadIdx.x >> 5));   // provably uniform (shared-memory addresses follow)
#else
    const int warp = threadIdx.x >> 5;
#endif
    const int K = P.K, K1 = K - 1, TS = P.TS;
    constexpr int KV = KPL * 32;
    constexpr int KVS = KV + 8;  // padded stride of the shared K-vectors (TS <= KV+4, block reads <= KV+6)

    // ---- shared memory carve-up (bfgs_smem_* in stm_b200.cu mirror this) -----------------------
    // [small block x nwarps][tile block x (nwarps - tm_warps)].  Warps < tm_warps keep their beta
    // tile in TENSOR MEMORY (tcgen05.st / tcgen05.ld, lane <-> word, columns <-> topics): TMEM is
    // 256 KB of otherwise idle on-chip storage, which more than doubles the documents in flight per SM.
    const int nwarps = blockDim.x >> 5;
    const bool is_tm = warp < P.tm_warps;
    unsigned char* small = smem_raw + (size_t)warp * P.smem_small;
    double* vec = reinterpret_cast<double*>(small);                       // [4][KVS]
#if STM_LS_REGS
    LsState S_regs;
    LsState& S = S_regs;
#else
    LsState& S = *reinterpret_cast<LsState*>(vec + 4 * KVS);
#endif
    uint64_t* mbar = reinterpret_cast<uint64_t*>(small + 4 * KVS * 8 + sizeof(LsState));
    unsigned char* tbase = smem_raw + (size_t)nwarps * P.smem_small +
                           (size_t)(is_tm ? 0 : warp - P.tm_warps) * P.smem_per_warp;
    const size_t tile_bytes = ((size_t)P.n_cap * TS * 4 + 127) & ~(size_t)127;
    float* tile = reinterpret_cast<float*>(tbase);                        // smem warps only
    float* cw = reinterpret_cast<float*>(tbase + tile_bytes);             // [n_cap] counts
    double* v0 = vec;            // e / broadcast scratch
    double* v1 = vec + KVS;
    double* v2 = vec + 2 * KVS;
    double* v3 = vec + 3 * KVS;
    // c_v / colsum_v of the a_k precompute: borrows v1..v3 when it fits, else its own block
    double* wv = (P.n_cap <= 3 * KVS) ? v1 : reinterpret_cast<double*>(cw + ((P.n_cap + 1) & ~1));
    for (int i = lane; i < 4 * KVS; i += 32) vec[i] = 0.0;  // pads stay zero for the whole kernel

    if (lane == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;
    // TMEM: one allocation of all 512 columns per CTA (1 CTA per SM); warp w owns lanes
    // 32*(w&3)..+31 (the only lanes it can address) and columns (w>>2)*tm_cols..+tm_cols-1.
    __shared__ uint32_t tm_base_s;
    if (P.tm_warps > 0) {
        if (warp == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
                smem_u32(&tm_base_s)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const uint32_t taddr =
        is_tm ? tm_base_s + ((uint32_t)((warp & 3) * 32) << 16) + (uint32_t)((warp >> 2) * P.tm_cols) : 0u;
    const int CS = (K + 1) & ~1;   // TMEM columns per word slot
#if STM_DBG_TIMING
    long long dbg_t[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    double* Hk = P.scratch + (size_t)gwarp * P.scratch_stride;  // BFGS inverse Hessian [K1][K1]

    // diagonal of siginv, lane-distributed
    double Sd[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        Sd[i] = (k < K1) ? P.prior[k] : 0.0;
    }

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = (int)atomicAdd(P.queue, 1u);
#if STM_UNIFORM_LS
        qi = __reduce_max_sync(STM_FULL, qi);   // lanes != 0 hold 0: a provably uniform broadcast (document index, length, loop bounds follow)
#else
        qi = __shfl_sync(STM_FULL, qi, 0);
#endif
        if (qi >= P.n_docs) break;
        STM_T(t_doc0);
        const int d = P.docs[qi];
        const long long p0 = P.doc_ptr[d];
        const int n = (int)(P.doc_ptr[d + 1] - p0);
        const int asp = P.aspect ? P.aspect[d] : 0;
        const float* beta_a = P.beta_t + (size_t)asp * P.V * TS;

        double x[KPL], mu[KPL], p[KPL], g[KPL], gt[KPL], xt[KPL], a[KPL], ex[KPL], xt2[KPL], gt2[KPL];
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = lane + 32 * i;
            x[i] = (k < K1) ? P.eta[(size_t)d * K1 + k] : 0.0;
            mu[i] = (k < K1) ? P.mu[(size_t)d * K1 + k] : 0.0;
            p[i] = 0.0; g[i] = 0.0; gt[i] = 0.0; xt[i] = 0.0; a[i] = 0.0; ex[i] = 0.0; xt2[i] = 0.0; gt2[i] = 0.0;
        }
        double nsum_l = 0.0;
        float cwr[J];   // TMEM path: counts of this lane's words (slot j <-> word lane + 32 j)
#pragma unroll
        for (int j = 0; j < J; ++j) cwr[j] = 0.f;
        if (is_tm) {
            // ---- TMEM path: global -> registers -> tcgen05.st, column sums on the way -----------
            int widr[J];
            double rv[J];
            const float4* rowp[J];
#pragma unroll
            for (int j = 0; j < J; ++j) {
                const int v = lane + 32 * j;
                const bool ok = v < n;
                widr[j] = ok ? P.word_id[p0 + v] : 0;
                cwr[j] = ok ? P.count[p0 + v] : 0.f;
                nsum_l += (double)cwr[j];
                rowp[j] = reinterpret_cast<const float4*>(beta_a + (size_t)widr[j] * TS);
                rv[j] = 0.0;   // running column sum, then c_v / colsum_v
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
                    tm_st8(taddr + (uint32_t)(j * CS + c0), b0[j], b1[j]);
                }
            }
            if ((CS - c0) & 4) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float4 b = __ldg(rowp[j] + (c0 >> 2));
                    rv[j] += beta_f2d(b.x); rv[j] += beta_f2d(b.y); rv[j] += beta_f2d(b.z); rv[j] += beta_f2d(b.w);
                    tm_st4(taddr + (uint32_t)(j * CS + c0), b);
                }
                c0 += 4;
            }
            if ((CS - c0) & 2) {
#pragma unroll
                for (int j = 0; j < J; ++j) {
                    const float2 b = __ldg(reinterpret_cast<const float2*>(rowp[j]) + (c0 >> 1));
                    rv[j] += beta_f2d(b.x); rv[j] += beta_f2d(b.y);
                    tm_st2(taddr + (uint32_t)(j * CS + c0), b);
                }
            }
            tm_wait_st();
#pragma unroll
            for (int j = 0; j < J; ++j) rv[j] = (double)cwr[j] / rv[j];
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
            // ---- smem path: stage counts, TMA-gather the beta rows ------------------------------
            fence_proxy_async();  // previous document's generic writes to this smem precede async writes
            __syncwarp();
            if (lane == 0) mbar_expect_tx(mbar, (uint32_t)(n * TS * 4));
            __syncwarp();
            for (int v = lane; v < n; v += 32) {
                const int w = P.word_id[p0 + v];
                const float c = P.count[p0 + v];
                cw[v] = c;
                nsum_l += (double)c;
                tma_row_g2s(tile + (size_t)v * TS, beta_a + (size_t)w * TS, (uint32_t)(TS * 4), mbar);
            }
            mbar_wait(mbar, parity);
            parity ^= 1;
            __syncwarp();
            // ---- a_k = sum_v beta_kv c_v / colsum_v   (eta-independent part of df, stm.py:954) ----
            for (int v = lane; v < n; v += 32) {
                const float4* row = reinterpret_cast<const float4*>(tile + (size_t)v * TS);
                double cs = 0.0;
                for (int q = 0; q < TS / 4; ++q) {
                    const float4 b = row[q];
                    cs += beta_f2d(b.x); cs += beta_f2d(b.y); cs += beta_f2d(b.z); cs += beta_f2d(b.w);
                }
                wv[v] = (double)cw[v] / cs;
            }
            __syncwarp();
            for (int v = 0; v < n; ++v) {
                const double r = wv[v];
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    const int k = lane + 32 * i;
                    if (k < K) a[i] += beta_f2d(tile[(size_t)v * TS + k]) * r;
                }
            }
            __syncwarp();
            // wv may alias v1..v3: restore the zero pads the K-vector reads rely on
            if (P.n_cap <= 3 * KVS) for (int i = lane; i < 3 * KVS; i += 32) v1[i] = 0.0;
            __syncwarp();
        }
        const double Nsum = warp_sum(nsum_l);           // np.sum(word_count)       stm.py:955
        const double Nint = (double)(long long)Nsum;    // int(np.sum(word_count))  stm.py:933

        // =======================================================================================
        // BFGS (scipy/optimize/_optimize.py:1345-1526) as a warp-uniform state machine with ONE
        // objective-evaluation site.  Every evaluation returns f, the gradient gt and dphi = gt.p.
        // =======================================================================================
        STM_T(t_bfgs0);
        STM_TACC(0, t_doc0, t_bfgs0);   // slot 0: gather + a_k
        int ls = LS_INIT, k_it = 0, warnflag = 0, nfev = 0, done = 0;
        const int maxiter = K1 * 200;
        double alpha = 0.0, f_eval = 0.0, dphi = 0.0;
        int have_cache = 0, have_cache2 = 0;
        // ex[] / scale_e (= N / sum exp) belong to memo entry ex_owner (0: xt, 1: xt2, -1: neither): the curvature
        // certificate wants theta at the base point of the new search, which is the accepted trial xt
        int ex_owner = -1;
        double scale_e = 0.0;
        // phi'(alpha) = g . p of the two memoised points: what SciPy's memoised gradient gives back for a repeated
        // trial point is the SAME value as the first time, so the value of the fresh evaluation is kept (and the
        // memo-hit path no longer runs a reduction); invalid after the direction p has changed
        double dphi2 = 0.0;
        int dphi_ok = 0, dphi2_ok = 0;
        if (STM_LS_REGS || lane == 0) {
            S.old_fval = 0.0; S.old_old_fval = 0.0; S.gnorm = 0.0; S.derphi0 = 0.0; S.f2 = 0.0;
            S.brackt = 0; S.stage = 1; S.w1_it = 0; S.w2_i = 0; S.z_i = 0; S.a_safe = 0.0;
        }
        __syncwarp();

        const double c1 = 1e-4, c2 = 0.9, xtol = 1e-14, stpmin = 1e-100, stpmax = 1e100;

        if (STM_DBG_SKIP_BFGS) done = 1;
        while (!done) {
            // ---------------- evaluate f, g at x + alpha p --------------------------------------
            STM_T(t_ev0);
            {
                double xn[KPL];
                bool same0 = have_cache, same1 = have_cache2;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    xn[i] = __dadd_rn(x[i], __dmul_rn(alpha, p[i]));
                    same0 = same0 && (xn[i] == xt[i]);
                    same1 = same1 && (xn[i] == xt2[i]);
                }
                // Memoisation.  SciPy's ScalarFunction re-uses f and g for a trial point identical to
                // the LAST one (scipy/_differentiable_functions.py:391-401); f is deterministic, so
                // re-using the last TWO points is value-identical and removes the A,B,A,B,... tail of
                // a collapsing dcsrch interval (~20 % of all evaluations).
                const bool hit0 = __all_sync(STM_FULL, same0);
                const bool hit1 = !hit0 && __all_sync(STM_FULL, same1);
#if STM_DBG_TIMING
                dbg_t[14] += 1; dbg_t[15] += (hit0 || hit1) ? 1 : 0;   // slots 14 / 15: steps, memo hits
#endif
                if (hit1) {
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const double tx = xt[i]; xt[i] = xt2[i]; xt2[i] = tx;
                        const double tg = gt[i]; gt[i] = gt2[i]; gt2[i] = tg;
                    }
                    const double tf = f_eval; f_eval = S.f2; S.f2 = tf;
                    const double td = dphi; dphi = dphi2; dphi2 = td;
                    const int tk = dphi_ok; dphi_ok = dphi2_ok; dphi2_ok = tk;
                    have_cache2 = have_cache;  // both valid after a swap
                    ex_owner = (ex_owner >= 0) ? 1 - ex_owner : -1;
                }
                if (!hit0 && !hit1) {
#pragma unroll
                    for (int i = 0; i < KPL; ++i) { xt2[i] = xt[i]; gt2[i] = gt[i]; }
                    S.f2 = f_eval;
                    dphi2 = dphi; dphi2_ok = dphi_ok;
                    have_cache2 = have_cache;
                    nfev++;
                    have_cache = 1;
                    ex_owner = 0;
                    double et[KPL];
                    double m = -INFINITY;
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int k = lane + 32 * i;
                        xt[i] = xn[i];
                        et[i] = (k < K1) ? xn[i] : ((k == K1) ? 0.0 : -INFINITY);
                        m = nanmax(m, et[i]);
                    }
                    m = warp_max(m);
                    STM_T(t_e1);
                    STM_TACC(8, t_ev0, t_e1);       // slot 8: memo check + max reduce
                    // red[]: cnt, ssum, quad, data, (S d - a).p, ex.p  — reduced together after the contraction
                    double red[6] = {0.0, 0.0, 0.0, 0.0, 0.0, 0.0};
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int k = lane + 32 * i;
                        ex[i] = exp_noinline(et[i] - m);
                        if (k < K) { if (et[i] == m) red[0] += 1.0; else red[1] += ex[i]; }
                        v0[k] = (k < K) ? ex[i] : 0.0;
                        const double dk = xt[i] - mu[i];
                        const double sdk = Sd[i] * dk;
                        red[2] += sdk * dk;
                        if (k < K1) { red[4] += (sdk - a[i]) * p[i]; red[5] += ex[i] * p[i]; }
                    }
                    __syncwarp();

                    STM_T(t_e2);
                    STM_TACC(9, t_e1, t_e2);        // slot 9: exp + partials + sync
                    // data term: sum_v c_v (m + log(sum_k e_k beta_kv))           stm.py:938-941
                    LogProd lp;
                    logprod_init(lp);
                    for (int w0 = 0; w0 < n; w0 += 32 * J) {
                        double acc[J][2];
                        float cj[J];
#pragma unroll
                        for (int j = 0; j < J; ++j) { acc[j][0] = 0.0; acc[j][1] = 0.0; }
                        if (is_tm) {
                            // tile in TMEM: lane <-> word (slot j), 8 topics per tcgen05.ld (one pass: n <= 32 J)
                            const double2* e2 = reinterpret_cast<const double2*>(v0);
                            int c0 = 0;
#pragma unroll 1
                            for (; c0 + 8 <= CS; c0 += 8) {
                                uint32_t b[J][8];
#pragma unroll
                                for (int j = 0; j < J; ++j) tm_ld8(taddr + (uint32_t)(j * CS + c0), b[j]);
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
                                for (int j = 0; j < J; ++j) tm_ld4(taddr + (uint32_t)(j * CS + c0), b[j]);
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
                                for (int j = 0; j < J; ++j) tm_ld2(taddr + (uint32_t)(j * CS + c0), b[j]);
                                const double2 e0 = e2[(c0 >> 1)];
                                tm_wait_ld();
#pragma unroll
                                for (int j = 0; j < J; ++j) {
                                    acc[j][0] = fma(e0.x, beta_u2d(b[j][0]), acc[j][0]);
                                    acc[j][1] = fma(e0.y, beta_u2d(b[j][1]), acc[j][1]);
                                }
                            }
#pragma unroll
                            for (int j = 0; j < J; ++j) cj[j] = cwr[j];
                        } else {
                            const float4* rows[J];
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                                            int v = w0 + lane + 32 * j;
                                if (v >= n) v = n - 1;
                                rows[j] = reinterpret_cast<const float4*>(tile + (size_t)v * TS);
                            }
                            const double2* e2 = reinterpret_cast<const double2*>(v0);
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
#pragma unroll
                            for (int j = 0; j < J; ++j) {
                                const int v = w0 + lane + 32 * j;
                                cj[j] = (v < n) ? cw[v] : 0.f;
                            }
                        }
#pragma unroll
                        for (int j = 0; j < J; ++j) {
                            const int v = w0 + lane + 32 * j;
                            if (v < n) logprod_add(lp, acc[j][0] + acc[j][1], cj[j]);
                        }
                        logprod_renorm(lp);
                    }
                    STM_T(t_e3);
                    STM_TACC(10, t_e2, t_e3);       // slot 10: contraction + logprod accumulate
                    red[3] = logprod_value(lp);
                    STM_T(t_e4);
                    STM_TACC(11, t_e3, t_e4);       // slot 11: log of the product
                    warp_sum_n<6>(red);
                    STM_T(t_e5);
                    STM_TACC(12, t_e4, t_e5);       // slot 12: batched reduction
                    const double cnt = red[0];
                    double ssum = red[1];
                    const double se_all = ssum + cnt;
                    if (ssum != 0.0 && cnt != 1.0) ssum = ddiv(ssum, cnt);
                    // scipy.special.logsumexp (scipy/special/_logsumexp.py:201-247)
                    const double lse = log_noinline(1.0 + ssum) + ((cnt == 1.0) ? 0.0 : log_noinline(cnt)) + m;
                    const double quad = 0.5 * red[2];
                    const double data = m * Nsum + red[3];
                    f_eval = quad - (data - Nint * lse);

                    // gradient stm.py:946-958 (beta NOT weighted by exp(eta): reference quirk)
                    const double scale = ddiv(Nsum, se_all);
                    scale_e = scale;
#pragma unroll
                    for (int i = 0; i < KPL; ++i) {
                        const int k = lane + 32 * i;
                        const double dk = xt[i] - mu[i];
                        gt[i] = (k < K1) ? (Sd[i] * dk - (a[i] - scale * ex[i])) : 0.0;
                    }
                    dphi = red[4] + scale * red[5];   // = gt . p
                    dphi_ok = 1;
                    STM_T(t_e6);
                    STM_TACC(13, t_e5, t_e6);       // slot 13: lse + f + gradient
                } else if (!dphi_ok) {
                    double dp_l2 = 0.0;
#pragma unroll
                    for (int i = 0; i < KPL; ++i) dp_l2 += gt[i] * p[i];
                    dphi = warp_sum(dp_l2);
                    dphi_ok = 1;
                }
            }

            // ---------------- consume the evaluation --------------------------------------------
#if STM_UNIFORM_LS
            f_eval = uniform_f64(f_eval);
            dphi = uniform_f64(dphi);
#endif
            STM_T(t_ev1);
            STM_TACC(1, t_ev0, t_ev1);      // slot 1: evaluation (incl. memo check)
            int accept = 0;      // 1: step alpha accepted (f_eval, gt valid)
            int fail = 0;        // 1: line search failed
            int new_iter = 0;    // 1: start a new BFGS iteration (compute p, init wolfe1)
            int start_w2 = 0;
            int start_zoom = 0;
            double zl = 0, zh = 0, zpl = 0, zph = 0, zdl = 0;

            if (ls == LS_INIT) {
                S.old_fval = f_eval;
                double n2 = 0.0, gm = 0.0;
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    g[i] = gt[i];
                    n2 += g[i] * g[i];
                    gm = nanmax(gm, fabs(g[i]));
                }
                S.old_old_fval = S.old_fval + dsqrt(warp_sum(n2)) * 0.5;
                S.gnorm = warp_max(gm);
                new_iter = 1;
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
#if STM_DCSTEP_INLINE
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
#else
                    if (S.stage == 1 && f <= S.fx && f > ftest) {
                        double fm = f - stp * S.gtest, fxm = S.fx - S.stx * S.gtest, fym = S.fy - S.sty * S.gtest;
                        double gm = gd - S.gtest, gxm = S.gx - S.gtest, gym = S.gy - S.gtest;
                        dcstep(S.stx, fxm, gxm, S.sty, fym, gym, stp, fm, gm, S.brackt, S.stmin, S.stmax);
                        S.fx = fxm + S.stx * S.gtest; S.fy = fym + S.sty * S.gtest;
                        S.gx = gxm + S.gtest; S.gy = gym + S.gtest;
                    } else {
                        dcstep(S.stx, S.fx, S.gx, S.sty, S.fy, S.gy, stp, f, gd, S.brackt, S.stmin, S.stmax);
                    }
#endif
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
                    bool dead = false;
#if STM_CURV_CERT
                    if (S.brackt && S.a_safe > 0.0 && S.stmax <= S.a_safe) dead = true;   // curvature certificate: cannot converge any more
#endif
                    if (dead || !isfinite(stp) || S.w1_it >= 100) start_w2 = 1;  // WARN / maxiter -> stp None
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
#if STM_CURV_CERT
            // curvature certificate: with a_lo <= a_hi _zoom's trials stay inside [a_lo, a_hi] (cubic / quadratic steps
            // are range-checked against margins of delta * (a_hi - a_lo) >= 0, else bisection — for a_hi < a_lo those
            // margins are negative and a trial may leave the interval, so that case is left to the replay), every
            // trial has phi' < 0, which keeps a_lo <= a_hi in both update branches (_linesearch.py:611-627), the
            // interval only shrinks, and no step in it can be accepted: _zoom runs out of iterations and returns None
            if (ls == LS_ZOOM && !accept && !fail && S.a_safe > 0.0 && S.a_lo >= 0.0 && S.a_lo <= S.a_hi && S.a_hi <= S.a_safe)
                fail = 1;
#endif
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

            if (fail) { warnflag = 2; done = 1; }
            STM_T(t_ls1);
            STM_TACC(2, t_ev1, t_ls1);      // slot 2: line-search logic

            if (accept) {
                // _minimize_bfgs body after the line search (scipy/optimize/_optimize.py:1452-1498)
                const double alpha_k = alpha;
                double ys_l = 0.0, gm = 0.0, pm = 0.0;
                double sk[KPL], yk[KPL];
#pragma unroll
                for (int i = 0; i < KPL; ++i) {
                    sk[i] = __dmul_rn(alpha_k, p[i]);
                    x[i] = xt[i];                 // xk + alpha_k*pk, same expression as the trial point
                    yk[i] = gt[i] - g[i];
                    g[i] = gt[i];
                    ys_l += yk[i] * sk[i];
                    gm = nanmax(gm, fabs(g[i]));
                    pm = nanmax(pm, fabs(p[i]));
                }
                S.old_old_fval = S.old_fval;
                S.old_fval = f_eval;
                k_it++;
                S.gnorm = warp_max(gm);
                pm = warp_max(pm);
                if (S.gnorm <= 1e-5) { done = 1; }
                else if (alpha_k * pm <= 0.0) { done = 1; }
                else if (!isfinite(S.old_fval)) { warnflag = 2; done = 1; }
                else {
                    const double rhok_inv = warp_sum(ys_l);
                    const double rho = (rhok_inv == 0.0) ? 1000.0 : ddiv(1.0, rhok_inv);
                    // Hk <- (I - rho s y')(Hk)(I - rho y s') + rho s s'   as a symmetric rank-2 update:
                    //   u = Hk y ;  Hk' = Hk - (rho u) s' - s (rho u)' + (rho^2 y'u + rho) s s'
                    double u[KPL];
                    if (k_it == 1) {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) u[i] = yk[i];  // Hk = I
                    } else {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) { u[i] = 0.0; v1[lane + 32 * i] = yk[i]; }
                        __syncwarp();
                        for (int r = 0; r < K1; ++r) {
                            const double yr = v1[r];
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
                        v1[k] = sk[i]; v2[k] = ru[i]; v3[k] = g[i];
                    }
                    __syncwarp();
                    // fused: write Hk' and accumulate p = -Hk' g
                    double pn[KPL];
#pragma unroll
                    for (int i = 0; i < KPL; ++i) pn[i] = 0.0;
                    for (int r = 0; r < K1; ++r) {
                        const double sr = v1[r], rur = v2[r], gr = v3[r];
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            if (k < K1) {
                                double h = (k_it == 1) ? ((r == k) ? 1.0 : 0.0) : Hk[(size_t)r * K1 + k];
                                h = h - (__dmul_rn(rur, sk[i]) + __dmul_rn(sr, ru[i])) + cc * sr * sk[i];
                                Hk[(size_t)r * K1 + k] = h;
                                pn[i] = fma(h, gr, pn[i]);
                            }
                        }
                    }
                    __syncwarp();
#pragma unroll
                    for (int i = 0; i < KPL; ++i) p[i] = -pn[i];
                    new_iter = 2;  // p already computed
                }
            }

            if (new_iter && !done) {
                if (!(S.gnorm > 1e-5) || !(k_it < maxiter)) {
                    done = 1;
                } else {
                    if (new_iter == 1) {
#pragma unroll
                        for (int i = 0; i < KPL; ++i) p[i] = -g[i];  // Hk = I
                    }
                    dphi_ok = 0; dphi2_ok = 0;   // new direction: the memoised g . p are stale
                    double d_l = 0.0;
#pragma unroll
                    for (int i = 0; i < KPL; ++i) d_l += g[i] * p[i];
                    S.derphi0 = warp_sum(d_l);
#if STM_CURV_CERT
                    // ---- curvature certificate -----------------------------------------------------------------
                    // The reference's gradient (stm.py:946-958) is NOT the gradient of its objective but of the convex
                    //   h(eta) = 1/2 (eta-mu)' S (eta-mu) - a' eta + N logsumexp([eta, 0])
                    // (SURVEY 8a, a6), so phi'(alpha) = g(x + alpha p) . p is non-decreasing in alpha and
                    //   phi'(alpha) - phi'(0) = int_0^alpha [ p'Sp + N Var_theta(t)(pt) ] dt,   pt = [p, 0],
                    // theta(t) = softmax([x + t p, 0]).  Two bounds on the variance term, valid for every t:
                    //   (i)  Var <= min(max_k pt_k^2, 1/2 |p|^2)             (diag(theta) - theta theta' <= diag(theta), norm <= 1/2)
                    //   (ii) Var_theta(t) <= e^{tR} Var_theta(0),  R = max pt - min pt   (theta_k(t) <= e^{tR} theta_k(0));
                    //        e^{tR} <= 1.11 while t R <= 0.1.
                    // Every acceptance test of the three searches contains the strong-Wolfe curvature condition
                    // |phi'(alpha)| <= 0.9 |phi'(0)| (_dcsrch.py:373, _linesearch.py:433, :606).  For alpha <= a_safe =
                    // 0.09 |phi'(0)| / C (C from (i), or from (ii) with a_safe R <= 0.1) the exact phi'(alpha) lies in
                    // [phi'(0), 0.91 phi'(0)], so the test fails by a margin of 0.01 |phi'(0)|.  The guard below bounds the
                    // rounding error of ANY fp64 evaluation of g . p by 1e-12 sum_i |p_i| (|S (x-mu)|_i + |a_i| + N) — three
                    // orders above the real thing — and gives no certificate (a_safe = 0) unless that is below the margin.
                    // Consequence: once a search's bracket lies inside [0, a_safe] and can only shrink (DCSRCH with
                    // brackt set; _zoom always), no later trial can be accepted and the search is known to fail —
                    // which is all that is left of it: a failed DCSRCH hands nothing to Wolfe-2, a failed _zoom raises
                    // _LineSearchError and BFGS returns the current x (_optimize.py:1446-1449).  The C oracle replays
                    // every search in full and checks this rule on the way (stm_oracle_shortcut_check): 0 acceptances
                    // after the certificate in every state of tests/ and tools/.
                    {
                        const bool th_ok = (ex_owner == 0);   // ex[], scale_e describe theta at x
                        double c_l = 0.0, s_l = 0.0, n_l = 0.0, m_l = 0.0, hi_l = 0.0, lo_l = 0.0;
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            const double pp = p[i] * p[i];
                            c_l += Sd[i] * pp; s_l += pp;
                            hi_l = fmax(hi_l, p[i]); lo_l = fmax(lo_l, -p[i]);
                            n_l += fabs(p[i]) * (fabs(Sd[i] * (x[i] - mu[i])) + fabs(a[i]) + Nsum);
                            if (k < K) m_l += scale_e * ex[i] * p[i];
                        }
                        const double pSp = warp_sum(c_l), pp2 = warp_sum(s_l), noise = 1e-12 * warp_sum(n_l);
                        const double m1 = ddiv(warp_sum(m_l), Nsum);          // mean of pt under theta(0)
                        const double phi = warp_max(hi_l), plo = warp_max(lo_l);   // max(pt), -min(pt)  (pt contains 0)
                        double v_l = 0.0;
#pragma unroll
                        for (int i = 0; i < KPL; ++i) {
                            const int k = lane + 32 * i;
                            const double dk = p[i] - m1;
                            if (k < K) v_l += scale_e * ex[i] * dk * dk;
                        }
                        const double NV0 = warp_sum(v_l);                       // N Var_theta(0)(pt)
                        const double pmax2 = fmax(phi * phi, plo * plo), R = phi + plo;
                        const double C1 = pSp + Nsum * fmin(pmax2, 0.5 * pp2);
                        double as = 0.0;
                        if (S.derphi0 < 0.0 && C1 > 0.0 && C1 < 1e300 && noise <= 0.01 * -S.derphi0) {
                            const double num = 0.09 * -S.derphi0;
                            as = ddiv(num, C1);
                            const double C2 = pSp + 1.1100001 * NV0;
                            if (th_ok && C2 > 0.0 && NV0 >= 0.0 && R > 0.0 && R < 1e300) {
                                const double as2 = py_min2(ddiv(num, C2), ddiv(0.1, R));
                                if (as2 > as) as = as2;
                            }
                        }
                        S.a_safe = isfinite(as) ? as : 0.0;
                    }
#endif
                    // scalar_search_wolfe1 prologue + DCSRCH START
                    double alpha1;
                    if (S.derphi0 != 0.0) {
                        alpha1 = py_min2(1.0, ddiv(1.01 * 2 * (S.old_fval - S.old_old_fval), S.derphi0));
                        if (alpha1 < 0.0) alpha1 = 1.0;
                    } else alpha1 = 1.0;
                    if (alpha1 < stpmin || alpha1 > stpmax || S.derphi0 >= 0.0 || !isfinite(alpha1)) {
                        // task = ERROR -> stp None -> wolfe2
                        double a1;
                        if (S.derphi0 != 0.0) a1 = py_min2(1.0, ddiv(1.01 * 2 * (S.old_fval - S.old_old_fval), S.derphi0));
                        else a1 = 1.0;
                        if (a1 < 0.0) a1 = 1.0;
                        a1 = py_min2(a1, 1e100);
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
#if STM_DBG_TIMING
            { const long long t_it1 = clock64(); dbg_t[3] += t_it1 - t_ls1; }  // slot 3: accept / H update / new iteration
#endif
        }  // BFGS loop
        STM_T(t_post0);

        int status = warnflag;
        if (status != 2) {
            if (k_it >= maxiter) status = 1;
            else {
                bool bad = isnan(S.gnorm) || isnan(S.old_fval);
#pragma unroll
                for (int i = 0; i < KPL; ++i) bad = bad || isnan(x[i]);
                status = __any_sync(STM_FULL, bad) ? 3 : 0;
            }
        }
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = lane + 32 * i;
            if (k < K1) P.eta[(size_t)d * K1 + k] = x[i];
        }
        if (lane == 0) {
            const int nit_c = k_it > 0xfffff ? 0xfffff : k_it;
            P.doc_info[d] = status | (nit_c << 4);
            P.doc_nfev[d] = nfev;
        }
        __syncwarp();
#if STM_DBG_TIMING
        { const long long t_p4 = clock64(); dbg_t[4] += t_p4 - t_post0; }   // slot 4: status + stores
#endif
    }  // document loop
#if STM_DBG_TIMING
    if (lane == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(P.dbg_cycles + i, (unsigned long long)dbg_t[i]);
#endif
    if (P.tm_warps > 0) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_s));
    }
}

// ===== POST KERNEL =====
// Kernel B: everything after the optimiser for one document (stm.py:546-590): theta, phi -> beta_ss,
// Hessian, make_pd / Cholesky, bound, nu -> sigma_ss.  Reads the eta kernel A wrote; ORs the repair
// stage into doc_info.
template <int KPL>
__global__ void __launch_bounds__(256, 1) post_kernel(const EstepParams P) {
    extern __shared__ __align__(128) unsigned char smem_raw[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int K = P.K, K1 = K - 1, TS = P.TS;
    constexpr int KV = KPL * 32;
    constexpr int KVS = KV + 8;  // padded stride of the shared K-vectors (TS <= KV+4, block reads <= KV+6)
    const int HS = K1 | 1;  // odd row stride (doubles) of the dense (K-1)x(K-1) matrices in smem

    // ---- per-warp shared memory carve-up (post_smem_per_warp in stm_b200.cu mirrors this) -----
    unsigned char* base = smem_raw + (size_t)warp * P.smem_per_warp;
    size_t tile_bytes = (size_t)P.n_cap * TS * 4;
    const size_t h_bytes = (size_t)K1 * HS * 8;
    if (h_bytes > tile_bytes) tile_bytes = h_bytes;
    tile_bytes = (tile_bytes + 127) & ~(size_t)127;
    float* tile = reinterpret_cast<float*>(base);
    double* Hm = reinterpret_cast<double*>(base);  // aliases the tile once it is dead
    double* wv = reinterpret_cast<double*>(base + tile_bytes);            // [n_cap] per-word fp64
    double* wv2 = wv + P.n_cap;                                           // [n_cap] sqrt(c_v)
    double* vec = wv2 + P.n_cap;                                          // [4][KVS]
    float* cw = reinterpret_cast<float*>(vec + 4 * KVS);                  // [n_cap] counts
    int* wid = reinterpret_cast<int*>(cw + P.n_cap);                      // [n_cap]
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wid + ((P.n_cap + 1) & ~1));
    double* v0 = vec;            // e / broadcast scratch
    double* v1 = vec + KVS;
    double* v2 = vec + 2 * KVS;
    double* v3 = vec + 3 * KVS;
    for (int i = lane; i < 4 * KVS; i += 32) vec[i] = 0.0;  // pads stay zero for the whole kernel

    if (lane == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    __syncwarp();
    uint32_t parity = 0;
#if STM_DBG_TIMING
    long long dbg_t[16] = {0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0, 0};
#endif

    const int gwarp = blockIdx.x * (blockDim.x >> 5) + warp;
    double* Hg = P.scratch + (size_t)gwarp * P.scratch_stride + (size_t)K1 * K1;  // Laplace Hessian bounce [K1][K1]
    double* sig_acc = P.sigma_ss_rep + (size_t)(gwarp % P.n_rep) * K1 * K1;

    // diagonal of siginv, lane-distributed
    double Sd[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        Sd[i] = (k < K1) ? P.prior[k] : 0.0;
    }
    const double sigmaentropy = P.prior[K1];

    for (;;) {
        int qi = 0;
        if (lane == 0) qi = (int)atomicAdd(P.queue, 1u);
        qi = __shfl_sync(STM_FULL, qi, 0);
        if (qi >= P.n_docs) break;
        STM_T(t_post0);
        const int d = P.docs[qi];
        const long long p0 = P.doc_ptr[d];
        const int n = (int)(P.doc_ptr[d + 1] - p0);
        const int asp = P.aspect ? P.aspect[d] : 0;
        const float* beta_a = P.beta_t + (size_t)asp * P.V * TS;

        // ---- stage ids / counts, then TMA-gather the beta rows --------------------------------
        fence_proxy_async();  // previous document's generic writes to this smem precede async writes
        __syncwarp();
        if (lane == 0) mbar_expect_tx(mbar, (uint32_t)(n * TS * 4));
        __syncwarp();
        double nsum_l = 0.0;
        for (int v = lane; v < n; v += 32) {
            const int w = P.word_id[p0 + v];
            const float c = P.count[p0 + v];
            wid[v] = w;
            cw[v] = c;
            nsum_l += (double)c;
            tma_row_g2s(tile + (size_t)v * TS, beta_a + (size_t)w * TS, (uint32_t)(TS * 4), mbar);
        }
        const double Nsum = warp_sum(nsum_l);           // np.sum(word_count)       stm.py:955

        double x[KPL], mu[KPL];
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = lane + 32 * i;
            x[i] = (k < K1) ? P.eta[(size_t)d * K1 + k] : 0.0;
            mu[i] = (k < K1) ? P.mu[(size_t)d * K1 + k] : 0.0;
        }
        mbar_wait(mbar, parity);
        parity ^= 1;
        __syncwarp();
        // =======================================================================================
        // post-optimisation: theta, phi -> beta_ss, Hessian, Cholesky, bound, nu -> sigma_ss
        // =======================================================================================
        double th[KPL], eu[KPL];  // stable softmax (stm.py:905-909), unshifted exp(eta~)
        {
            double et[KPL], m = -INFINITY, se_l = 0.0, ss_l = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                et[i] = (k < K1) ? x[i] : ((k == K1) ? 0.0 : -INFINITY);
                m = nanmax(m, et[i]);
            }
            m = warp_max(m);
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                eu[i] = exp(et[i]);
                th[i] = exp(et[i] - m);
                se_l += eu[i];
                ss_l += th[i];
            }
            const double se = warp_sum(se_l), ss = warp_sum(ss_l);
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                th[i] = th[i] / ss;
                if (k < K) P.theta[(size_t)d * K + k] = eu[i] / se;  // stm.py:547-549 (no max shift)
                v0[k] = (k < K) ? eu[i] : 0.0;
                v1[k] = (k < K) ? th[i] * eu[i] : 0.0;
            }
            __syncwarp();
        }
        if (STM_DBG_SKIP_POST) { if (lane == 0) P.doc_bound[d] = 0.0; continue; }
        // colsum_v = sum_k e_k beta_kv and the log-likelihood part of the bound (stm.py:1088-1096)
        STM_T(t_p1);
        STM_TACC(4, t_post0, t_p1);         // slot 4: status, theta
        double loglik;
        {
            LogProd lp;
            logprod_init(lp);
            const double2* e2 = reinterpret_cast<const double2*>(v0);
            const double2* t2 = reinterpret_cast<const double2*>(v1);
#pragma unroll 1
            for (int v = lane; v < n; v += 32) {
                const float4* row = reinterpret_cast<const float4*>(tile + (size_t)v * TS);
                double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll 1
                for (int q = 0; q < TS / 4; ++q) {
                    const float4 bq = row[q];
                    const double2 ea = e2[2 * q], eb = e2[2 * q + 1], wa = t2[2 * q], wb = t2[2 * q + 1];
                    const double b0 = beta_f2d(bq.x), b1 = beta_f2d(bq.y), b2 = beta_f2d(bq.z), b3 = beta_f2d(bq.w);
                    s0 = fma(ea.x, b0, s0); t0 = fma(wa.x, b0, t0);
                    s1 = fma(ea.y, b1, s1); t1 = fma(wa.y, b1, t1);
                    s0 = fma(eb.x, b2, s0); t0 = fma(wb.x, b2, t0);
                    s1 = fma(eb.y, b3, s1); t1 = fma(wb.y, b3, t1);
                }
                const float cf = cw[v];
                logprod_add(lp, t0 + t1, cf);
                logprod_renorm(lp);
                const double sq = sqrt((double)cf);
                wv[v] = sq / (s0 + s1);   // sqrt(c_v)/colsum_v
                wv2[v] = sq;
            }
            loglik = warp_sum(logprod_value(lp));
        }
        __syncwarp();

        // Hessian data term  sum_v b_v b_v'  with b_kv = beta_kv e_k sqrt(c_v)/colsum_v (stm.py:1000-1006)
        // on the fp64 tensor cores: mma.sync m8n8k4 (DMMA), four words per step.  Lane (w = lane&3,
        // kk = lane>>2) computes b for word v0+w and topics 8t+kk: exactly its A-fragment element of
        // row-block t AND its B-fragment element of column-block t, so no staging and no barrier.
        // phi_kv = b_kv sqrt(c_v) is scattered into beta_ss on the way (stm.py:1103-1118, 582-590).
        double rowsum[KPL];  // sum_v phi_kv  (np.sum(c, axis=1), stm.py:1011)
        {
            constexpr int NBMAX = 4 * KPL;                 // 8-topic blocks covering K <= 32*KPL
            constexpr int RB = (KPL <= 2) ? NBMAX : ((KPL == 3) ? 3 : 2);  // block rows per pass
            const int nb = (K1 + 7) >> 3;                  // block rows/cols of the (K-1)x(K-1) matrix
            const int w4 = lane & 3, kk = lane >> 2;
            double* beta_ss_a = P.beta_ss_t + (size_t)asp * P.V * TS;
            double rs[NBMAX];
#pragma unroll
            for (int t = 0; t < NBMAX; ++t) rs[t] = 0.0;
#pragma unroll 1
            for (int r0 = 0; r0 < nb; r0 += RB) {
                double acc[RB][NBMAX][2];
#pragma unroll
                for (int rr = 0; rr < RB; ++rr)
#pragma unroll
                    for (int bc = 0; bc < NBMAX; ++bc) { acc[rr][bc][0] = 0.0; acc[rr][bc][1] = 0.0; }
#pragma unroll 1
                for (int vb = 0; vb < n; vb += 4) {
                    const int v = vb + w4;
                    const bool vok = v < n;
                    const int vc = vok ? v : 0;
                    const double sc = vok ? wv[vc] : 0.0;
                    const double sqc = wv2[vc];
                    const float* trow = tile + (size_t)vc * TS;
                    double* ssrow = beta_ss_a + (size_t)wid[vc] * TS;
                    double fr[NBMAX];
#pragma unroll
                    for (int t = 0; t < NBMAX; ++t) {
                        const int k = 8 * t + kk;
                        double bk = 0.0;
                        if (k < K) {
                            bk = (beta_f2d(trow[k]) * v0[k]) * sc;
                            if (r0 == 0 && vok) {
                                const double ph = bk * sqc;
                                rs[t] += ph;
                                if (!STM_DBG_NO_PHI) red_add_f64(ssrow + k, ph);
                            }
                        }
                        fr[t] = bk;
                    }
#pragma unroll
                    for (int rr = 0; rr < RB; ++rr) {
                        const int br = r0 + rr;
#pragma unroll
                        for (int bc = 0; bc < NBMAX; ++bc) {
                            // single pass (KPL <= 2): br == rr at compile time, the upper blocks vanish
                            if ((KPL <= 2) ? (bc <= rr) : true) {
                                if (br < nb && bc <= br) {
                                    const double af = (KPL <= 2) ? fr[rr] : fr_sel<NBMAX>(fr, br);
                                    asm volatile(
                                        "mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                                        : "+d"(acc[rr][bc][0]), "+d"(acc[rr][bc][1])
                                        : "d"(af), "d"(fr[bc]));
                                }
                            }
                        }
                    }
                }
                // C fragment: row = lane>>2, cols = 2*(lane&3) + {0,1}
#pragma unroll
                for (int rr = 0; rr < RB; ++rr) {
                    const int br = r0 + rr;
#pragma unroll
                    for (int bc = 0; bc < NBMAX; ++bc) {
                        if ((KPL <= 2) ? (bc <= rr) : true) {
                            if (br < nb && bc <= br) {
                                const int gi = br * 8 + kk;
#pragma unroll
                                for (int c = 0; c < 2; ++c) {
                                    const int gj = bc * 8 + 2 * w4 + c;
                                    if (gi < K1 && gj <= gi) Hg[(size_t)gi * K1 + gj] = acc[rr][bc][c];
                                }
                            }
                        }
                    }
                }
            }
            // rowsum_k: add the 4 word slots (lanes sharing kk), publish, re-read lane-distributed
#pragma unroll
            for (int t = 0; t < NBMAX; ++t) {
                rs[t] += __shfl_xor_sync(STM_FULL, rs[t], 1);
                rs[t] += __shfl_xor_sync(STM_FULL, rs[t], 2);
            }
            __syncwarp();
#pragma unroll
            for (int t = 0; t < NBMAX; ++t)
                if (w4 == 0 && 8 * t + kk < KV) v1[8 * t + kk] = rs[t];
            __syncwarp();
#pragma unroll
            for (int i = 0; i < KPL; ++i) rowsum[i] = v1[lane + 32 * i];
        }
        __syncwarp();
        __threadfence_block();
        // assemble H = data - N theta theta' + diag(-rowsum + N theta) + siginv into smem (tile is dead)
#pragma unroll
        for (int i = 0; i < KPL; ++i) {
            const int k = lane + 32 * i;
            v0[k] = th[i]; v1[k] = rowsum[i];
        }
        __syncwarp();
        for (int r = 0; r < K1; ++r) {
            const double thr = v0[r];
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                if (k <= r) {
                    double h = __ldcg(&Hg[(size_t)r * K1 + k]) - Nsum * (thr * th[i]);
                    if (k == r) h = (h - rowsum[i] + Nsum * th[i]) + Sd[i];
                    Hm[(size_t)r * HS + k] = h;
                    Hm[(size_t)k * HS + r] = h;
                }
            }
        }
        __syncwarp();

        // PD test + repairs (stm.py:1017-1021, 1039-1048).  "all eigenvalues > 0" is restated as
        // "Cholesky succeeds".  The factor overwrites the strict lower triangle + a separate diagonal
        // so the matrix can be restored from its upper triangle for the retries.
        if (STM_DBG_SKIP_DENSE) { if (lane == 0) P.doc_bound[d] = 0.0; continue; }
        STM_T(t_p2);
        STM_TACC(5, t_p1, t_p2);            // slot 5: colsum + Hessian (DMMA) + assemble
        int repair = 0;
        double* Ld = v2;      // diagonal of L
        double* Dg = v3;      // diagonal of H (original / repaired)
        double* Dw = v0;      // running diagonal  H_jj - sum_k<j L_jk^2  (same order as a sequential loop)
        for (int k = lane; k < K1; k += 32) Dg[k] = Hm[(size_t)k * HS + k];
        __syncwarp();
        int upper = 0;
        for (int attempt = 0;; ++attempt) {
            const int ok = chol_factor(Hm, Dw, Dg, Ld, K1, HS, lane);
            if (ok) break;
            // failed: restore the lower triangle from the upper one, then repair
            __syncwarp();
            for (int r = 0; r < K1; ++r)
                for (int k = lane; k < r; k += 32) Hm[(size_t)r * HS + k] = Hm[(size_t)k * HS + r];
            __syncwarp();
            if (attempt == 0 || attempt == 2 || attempt == 3) {
                // make_pd (stm.py:964-984): d_i <- max(d_i, sum_j!=i |H_ij|)
                for (int i = lane; i < K1; i += 32) {
                    double tot = 0.0;
                    for (int k = 0; k < K1; ++k) if (k != i) tot += fabs(Hm[(size_t)i * HS + k]);
                    tot += fabs(Dg[i]);
                    const double mag = tot - fabs(Dg[i]);
                    if (Dg[i] < mag) Dg[i] = mag;
                }
            }
            if (attempt == 1 || attempt == 3) {
                for (int i = lane; i < K1; i += 32) Dg[i] += 1e-5;
            }
            if (attempt == 0) repair = 1;            // hessian(): make_pd
            else if (attempt == 1) repair = 2;       // hessian(): + 1e-5
            else if (attempt == 2) repair += 4;      // decompose_hessian(): make_pd
            else if (attempt == 3) { repair += 8; upper = 1; }  // scipy.linalg.cholesky (upper) of make_pd + 1e-5 I
            else {
                for (int k = lane; k < K1; k += 32) Ld[k] = nan("");
                __syncwarp();
                break;
            }
            __syncwarp();
        }
        __syncwarp();

        STM_T(t_p3);
        STM_TACC(6, t_p2, t_p3);            // slot 6: Cholesky + repairs
        // bound (stm.py:1085-1100)
        double det_l = 0.0, q_l = 0.0;
        for (int k = lane; k < K1; k += 32) det_l += log_noinline(Ld[k]);
#pragma unroll
        for (int i = 0; i < KPL; ++i) { const double dk = x[i] - mu[i]; q_l += (Sd[i] * dk) * dk; }
        const double bound = loglik - warp_sum(det_l) - 0.5 * warp_sum(q_l) - sigmaentropy;
        if (lane == 0) {
            P.doc_bound[d] = bound;
            P.doc_info[d] = (P.doc_info[d] & 0xffffff) | (repair << 24);
        }

        // nu = H^-1 = L^-T L^-1 (stm.py:1052-1066), accumulated into sigma_ss (stm.py:582)
        if (!upper) {
            inverse_and_nu(Hm, Ld, sig_acc, K1, HS, lane);
        } else {
            for (int k = lane; k < K1; k += 32) {
                const double il = 1.0 / Ld[k];
                red_add_f64(sig_acc + (size_t)k * K1 + k, il * il);
            }
        }
        __syncwarp();
#if STM_DBG_TIMING
        { const long long t_p4 = clock64(); dbg_t[7] += t_p4 - t_p3; }   // slot 7: bound + inverse + nu
#endif
    }  // document loop
#if STM_DBG_TIMING
    if (lane == 0)
        for (int i = 0; i < 16; ++i) atomicAdd(P.dbg_cycles + i, (unsigned long long)dbg_t[i]);
#endif
}


// ===== POST KERNEL, GROUP VERSION =====
// Kernel B for K-1 <= 52: THREE warps (96 threads) per document, up to 6 documents per CTA, so the
// post-optimisation work runs at 18 warps per SM instead of 6 and its dense part is parallel:
//   * Hessian data term: DMMA row-blocks dealt to the 3 warps (snake order, balanced), phi scattered
//     and row sums taken by the warp that owns the row block;
//   * PD test, Cholesky pivots, nu = H^-1 in ONE pass: the symmetric SWEEP operator
//     (B_kk = -1/A_kk, B_ik = A_ik/A_kk, B_ij = A_ij - A_ik A_kj/A_kk; after all pivots B = -A^-1),
//     each thread holding a 4x4 patch of the lower triangle in REGISTERS (13*14/2 = 91 patches);
//     pivot k equals the Cholesky pivot L_kk^2, so "np.linalg.cholesky succeeds" (stm.py:1017, 1040)
//     is "all pivots > 0" and sum log L_ii = 1/2 sum log d_k;
//   * nu leaves the registers as fp64 reductions into the replicated sigma_ss accumulators.
// Group geometry.  GW warps (GT = 32 GW threads) per document, one 4x4 patch of the lower triangle per thread:
// nb4 = ceil((K-1)/4) block rows -> nb4 (nb4+1)/2 patches.  K-1 <= 52: 91 patches, GW = 3 (the C3 shape);
// K-1 <= 63: 136 patches, GW = 5; K <= 96: 300 patches, GW = 10; K-1 <= 100: 325, GW = 11; K <= 128: 528, GW = 17.
constexpr int POST_GW = 3;              // warps per document of the base configuration (host code: stm::POST_GT)
constexpr int POST_GT = POST_GW * 32;
__host__ __device__ constexpr int post_ust(int GW) {        // stride of the small fp64 vectors (Dg, u, d): >= 4 nb4
    return GW == 3 ? 56 : (GW == 5 ? 72 : (GW <= 11 ? 104 : 136));
}
#ifndef STM_HESS_FUSED
#define STM_HESS_FUSED 1     // kernel B, K = 49..56: a warp's block rows of the DMMA Hessian pass fused into one pass over the words
#endif
#ifndef STM_POST_TMEM
#define STM_POST_TMEM 1      // kernel B, K <= 64: the Hessian data term waits in TENSOR MEMORY (not in the L2 scratch) for the tile to die
#endif
#ifndef STM_ASM_BATCH
#define STM_ASM_BATCH 2      // (r02 A/B at C3 with stepped indices: 1 / 2 / 4 -> 8.34 / 8.32 / 8.37 ms) Hessian elements per thread whose L2 loads are in flight together (kernel B assembly)
#endif
#ifndef STM_HESS_UNROLL
#define STM_HESS_UNROLL 1    // unroll factor of the DMMA row pass over word quadruples
#endif
#ifndef STM_POST_MAX_THREADS
#define STM_POST_MAX_THREADS 576        // 6 groups of 96 threads -> 112 registers per thread
#endif
__host__ __device__ constexpr int post_group_max_threads(int GW) {   // launch bound: groups x GT
    return GW == 3 ? STM_POST_MAX_THREADS : (GW == 5 ? 480 : GW * 32);
}
__host__ __device__ constexpr int post_group_min_blocks(int GW) {    // CTAs per SM the register budget is cut for
    return (GW == 10 || GW == 11) ? 2 : 1;
}
constexpr int POST_RED = 24;            // doubles of per-group reduction scratch (partials 0..GW-1, scalars 20..)

template <int GW>
__device__ __forceinline__ void group_bar(int grp) {
    asm volatile("bar.sync %0, %1;" ::"r"(grp + 1), "r"(GW * 32) : "memory");
}
// sum over the threads of a group, partials added in warp order; red = group-private shared memory
template <int GW>
__device__ __forceinline__ double group_sum(double v, double* red, int wg, int lane, int grp) {
    v = warp_sum(v);
    if (lane == 0) red[wg] = v;
    group_bar<GW>(grp);
    double r = (red[0] + red[1]) + red[2];
#pragma unroll
    for (int w = 3; w < GW; ++w) r += red[w];
    group_bar<GW>(grp);
    return r;
}

// One row block (8 topics, BR static) of the Hessian data term over all words of a document, on the
// fp64 tensor cores: mma.sync m8n8k4, four words per step.  Lane (w4 = lane&3, kk = lane>>2) computes
// b = beta e sqrt(c)/colsum for word vb+w4 and topic 8t+kk: exactly its A-fragment element of row
// block t AND its B-fragment element of column block t.  The owner of the row block also scatters
// phi = b sqrt(c) into beta_ss and takes its row sums.
template <int NBMAX, int BR>
__device__ __forceinline__ void hess_row_pass(const float* tile, int TS, int n, const double* wv, const double* wv2,
                                              const int* wid, const double (&ek)[NBMAX], double ekb, bool kok,
                                              double* ssb, int w4, int kk, double (&acc)[NBMAX][2], double& rs) {
STM_PRAGMA_(unroll STM_HESS_UNROLL)
    for (int vb = 0; vb < n; vb += 4) {
        const int v = vb + w4;
        const double sc = wv[v];                      // zero for the (< 4) slots past the last word
        const float* tb = tile + (size_t)min(v, n - 1) * TS + kk;
        // sum_v b_v b_v' with b = beta e sqrt(c) / colsum = (beta e) sc: the word's scale goes into the A fragment as
        // sc^2, so that a B fragment costs one multiplication instead of two
        double fr[BR + 1];
#pragma unroll
        for (int t = 0; t <= BR; ++t) fr[t] = beta_f2d(tb[8 * t]) * ek[t];
        const double be = beta_f2d(tb[8 * BR]) * ekb;
        const double fa = be * (sc * sc);
        if (v < n && kok) {
            const double ph = be * (sc * wv2[v]);
            rs += ph;
            if (!STM_DBG_NO_PHI) red_add_f64(ssb + wid[v], ph);
        }
#pragma unroll
        for (int t = 0; t <= BR; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[t][0]), "+d"(acc[t][1])
                         : "d"(fa), "d"(fr[t]));
    }
}

// K = 49 .. 56 with three warps per document (the benchmark shape): a warp's two or three block rows (R0 > R1 > R2,
// -1 = none) in ONE pass over the words.  The B fragments fr[0 .. R0] are formed once per word quadruple and serve all
// of the warp's rows (separate passes form sum (R + 2) of them); every element sees the same operations in the same
// order as in hess_row_pass, so the sums are bit-identical.
template <int R0, int R1, int R2>
__device__ __forceinline__ void hess_rows_fused(const float* tile, int TS, int n, const double* wv, const double* wv2,
                                                const int* wid, const double (&ek)[8], int K, double* beta_ss_a, int w4,
                                                int kk, double (&acc0)[8][2], double (&acc1)[8][2],
                                                double (&acc2)[8][2], double& rs0, double& rs1, double& rs2) {
    const bool ok0 = 8 * R0 + kk < K, ok1 = R1 >= 0 && 8 * R1 + kk < K, ok2 = R2 >= 0 && 8 * R2 + kk < K;
    double* s0 = beta_ss_a + 8 * R0 + kk;
    double* s1 = beta_ss_a + 8 * (R1 < 0 ? 0 : R1) + kk;
    double* s2 = beta_ss_a + 8 * (R2 < 0 ? 0 : R2) + kk;
#pragma unroll 1
    for (int vb = 0; vb < n; vb += 4) {
        const int v = vb + w4;
        const double sc = wv[v];                      // zero for the (< 4) slots past the last word
        const float* tb = tile + (size_t)min(v, n - 1) * TS + kk;
        double fr[R0 + 1];
#pragma unroll
        for (int t = 0; t <= R0; ++t) fr[t] = beta_f2d(tb[8 * t]) * ek[t];
        const bool vin = v < n;
        const double sc2 = sc * sc;
        {
            const double fa = fr[R0] * sc2;
            if (vin && ok0) {
                const double ph = fr[R0] * (sc * wv2[v]);
                rs0 += ph;
                if (!STM_DBG_NO_PHI) red_add_f64(s0 + wid[v], ph);
            }
#pragma unroll
            for (int t = 0; t <= R0; ++t)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc0[t][0]), "+d"(acc0[t][1])
                             : "d"(fa), "d"(fr[t]));
        }
        if constexpr (R1 >= 0) {
            const double fa = fr[R1] * sc2;
            if (vin && ok1) {
                const double ph = fr[R1] * (sc * wv2[v]);
                rs1 += ph;
                if (!STM_DBG_NO_PHI) red_add_f64(s1 + wid[v], ph);
            }
#pragma unroll
            for (int t = 0; t <= R1; ++t)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc1[t][0]), "+d"(acc1[t][1])
                             : "d"(fa), "d"(fr[t]));
        }
        if constexpr (R2 >= 0) {
            const double fa = fr[R2] * sc2;
            if (vin && ok2) {
                const double ph = fr[R2] * (sc * wv2[v]);
                rs2 += ph;
                if (!STM_DBG_NO_PHI) red_add_f64(s2 + wid[v], ph);
            }
#pragma unroll
            for (int t = 0; t <= R2; ++t)
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                             : "+d"(acc2[t][0]), "+d"(acc2[t][1])
                             : "d"(fa), "d"(fr[t]));
        }
    }
}

// row-block dispatch for more than 8 block rows (K > 64): compile-time recursion instead of a switch
template <int NBMAX, int BR>
struct HessDispatch {
    static __device__ __forceinline__ void run(int br, const float* tile, int TS, int n, const double* wv,
                                               const double* wv2, const int* wid, const double (&ek)[NBMAX], double ekb,
                                               bool kok, double* ssb, int w4, int kk, double (&acc)[NBMAX][2], double& rs) {
        if (br == BR) hess_row_pass<NBMAX, BR>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
        else HessDispatch<NBMAX, BR - 1>::run(br, tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
    }
};
template <int NBMAX>
struct HessDispatch<NBMAX, -1> {
    static __device__ __forceinline__ void run(int, const float*, int, int, const double*, const double*, const int*,
                                               const double (&)[NBMAX], double, bool, double*, int, int,
                                               double (&)[NBMAX][2], double&) {}
};

// K > 64 (more than 8 block rows): the Hessian data term is cut into WORK UNITS (row block br, NC <= 7 consecutive
// column blocks from c0) instead of whole row blocks.  A unit needs 2 NC accumulators and NC exponentials instead of
// 2 (br + 1) and 16 (the 11- and 17-warp instantiations spilled ~1 KB per thread), and 19 units of <= 7 block pairs
// deal over 11 warps with a maximum of 9 pairs per warp where whole row blocks gave 13 (K = 100).  Static NC, run-time
// br / c0: 7 instantiations instead of one per block row.
template <int NC>
__device__ __forceinline__ void hess_unit_pass(const float* tile, int TS, int n, const double* wv, const double* wv2,
                                               const int* wid, const double (&ek)[7], double ekb, bool kok, double* ssb,
                                               int w4, int kk, int br, int c0, bool do_phi, double (&acc)[7][2],
                                               double& rs) {
#pragma unroll 1
    for (int vb = 0; vb < n; vb += 4) {
        const int v = vb + w4;
        const double sc = wv[v];                      // zero for the (< 4) slots past the last word
        const float* tb = tile + (size_t)min(v, n - 1) * TS + kk;
        double fr[NC];
#pragma unroll
        for (int t = 0; t < NC; ++t) fr[t] = beta_f2d(tb[8 * (c0 + t)]) * ek[t];
        const double be = beta_f2d(tb[8 * br]) * ekb;
        const double fa = be * (sc * sc);              // the word's scale enters once, squared, on the A side
        if (do_phi && v < n && kok) {
            const double ph = be * (sc * wv2[v]);
            rs += ph;
            if (!STM_DBG_NO_PHI) red_add_f64(ssb + wid[v], ph);
        }
#pragma unroll
        for (int t = 0; t < NC; ++t)
            asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
                         : "+d"(acc[t][0]), "+d"(acc[t][1])
                         : "d"(fa), "d"(fr[t]));
    }
}

template <int KPL, int GW>
__global__ void __launch_bounds__(post_group_max_threads(GW), post_group_min_blocks(GW)) post_group_kernel(const EstepParams P) {
    constexpr int POST_GW = GW, POST_GT = GW * 32, POST_UST = post_ust(GW);
    extern __shared__ __align__(128) unsigned char smem_raw[];
    // warp-uniform indices go through redux.sync so that the compiler KNOWS they are uniform (no
    // reconvergence scaffolding around the warp-synchronous instructions they guard)
    const int warp_u = __reduce_max_sync(STM_FULL, (int)(threadIdx.x >> 5));
    const int grp = warp_u / POST_GW;
    const int wg = warp_u - grp * POST_GW;
    const int lane = threadIdx.x & 31;
    const int gt = wg * 32 + lane;
    const int K = P.K, K1 = K - 1, TS = P.TS;
    constexpr int KV = KPL * 32;
    constexpr int KVS = KV + 8;
    constexpr int NBMAX = 4 * KPL;
    const int HS = K1 | 1;

    // ---- per-group shared memory carve-up (post_group_smem in stm_b200.cu mirrors this) --------
    unsigned char* base = smem_raw + (size_t)grp * P.smem_per_warp;
    size_t tile_bytes = (size_t)(P.n_cap + 1) * TS * 4;   // + 1 row: the last 8-topic block may read past a row's end
    const size_t h_bytes = (size_t)K1 * HS * 8;
    if (h_bytes > tile_bytes) tile_bytes = h_bytes;
    tile_bytes = (tile_bytes + 127) & ~(size_t)127;
    float* tile = reinterpret_cast<float*>(base);
    double* Hm = reinterpret_cast<double*>(base);   // aliases the tile once it is dead
    size_t w_bytes = (size_t)(P.n_cap + 4) * 16;
    if (w_bytes < (size_t)4 * POST_UST * 8) w_bytes = (size_t)4 * POST_UST * 8;
    double* wv = reinterpret_cast<double*>(base + tile_bytes);   // [n_cap + 4] sqrt(c_v)/colsum_v, zero past n
    double* wv2 = wv + P.n_cap + 4;                              // [n_cap + 4] sqrt(c_v)
    double* Dg = wv;                      // the dense phase re-uses the per-word block:
    double* ubuf = wv + POST_UST;         //   diagonal, 2 pivot vectors, pivots
    double* dvec = wv + 3 * POST_UST;
    double* vec = reinterpret_cast<double*>(base + tile_bytes + w_bytes);   // [4][KVS]
    double* red = vec + 4 * KVS;                                            // [POST_RED] reductions / scalars
    float* cw = reinterpret_cast<float*>(red + POST_RED);
    int* wid = reinterpret_cast<int*>(cw + P.n_cap);
    uint64_t* mbar = reinterpret_cast<uint64_t*>(wid + ((P.n_cap + 1) & ~1));
    int* qslot = reinterpret_cast<int*>(mbar + 1);
    double* v0 = vec;            // exp(eta~)
    double* v1 = vec + KVS;      // theta * exp(eta~)
    double* v2 = vec + 2 * KVS;  // theta (stable softmax)
    double* v3 = vec + 3 * KVS;  // row sums of phi
    for (int i = gt; i < 4 * KVS; i += POST_GT) vec[i] = 0.0;
    if (gt == 0) mbar_init(mbar, 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    group_bar<GW>(grp);
    uint32_t parity = 0;
    // K > 64: Hessian work units (block row, <= 7 consecutive column blocks), largest first (see hess_unit_pass)
    __shared__ unsigned short units_s[64];
    __shared__ int unit_count_s;
    if constexpr (NBMAX > 8) {
        if (threadIdx.x == 0) {
            const int nbp_ = (K + 7) >> 3;
            int cnt = 0;
            for (int sz = 7; sz >= 1; --sz)
                for (int br = nbp_ - 1; br >= 0; --br)
                    for (int j = 0; 7 * j <= br; ++j) {
                        const int nc = (br + 1 - 7 * j) < 7 ? (br + 1 - 7 * j) : 7;
                        if (nc == sz) units_s[cnt++] = (unsigned short)(br | (j << 4) | (nc << 8));
                    }
            unit_count_s = cnt;
        }
        __syncthreads();
    }

    // K <= 64: the DMMA accumulators of a warp's row passes wait in tensor memory (96 columns of the warp's own 32
    // lanes: 3 passes x 8 blocks x 2 doubles) until every warp of the group is done with the tile and H can be written
    // over it — kernel B has no other use for the 256 KB, and the L2 round trip of a bounce buffer was 17 % of its
    // stall samples.  One CTA per SM (registers), <= 20 warps: 5 x 96 columns per lane quadrant.
    constexpr bool USE_TM = (STM_POST_TMEM != 0) && NBMAX <= 8;
    __shared__ uint32_t tm_base_b;
    if constexpr (USE_TM) {
        if (warp_u == 0) {
            asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(&tm_base_b)));
            asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
        }
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        asm volatile("tcgen05.fence::after_thread_sync;");
    }
    const uint32_t tmw = USE_TM ? tm_base_b + ((uint32_t)((warp_u & 3) * 32) << 16) + (uint32_t)((warp_u >> 2) * 96) : 0u;

    const int ggrp = blockIdx.x * (blockDim.x / POST_GT) + grp;
    double* Hg = P.scratch + (size_t)ggrp * P.scratch_stride + (size_t)K1 * K1;   // Hessian bounce (L2; K > 64 only)
    double* sig_acc = P.sigma_ss_rep + (size_t)(ggrp % P.n_rep) * K1 * K1;

    // this thread's 4x4 patch of the lower triangle
    const int nb4 = (K1 + 3) >> 2;
    const bool has_patch = gt < nb4 * (nb4 + 1) / 2;
    int pi = 0, pj = 0;
    if (has_patch) {
        while ((pi + 1) * (pi + 2) / 2 <= gt) pi++;
        pj = gt - pi * (pi + 1) / 2;
    }
    double Sd[KPL];
#pragma unroll
    for (int i = 0; i < KPL; ++i) {
        const int k = lane + 32 * i;
        Sd[i] = (k < K1) ? P.prior[k] : 0.0;
    }
    const double sigmaentropy = P.prior[K1];

    for (;;) {
        if (gt == 0) *qslot = (int)atomicAdd(P.queue, 1u);
        fence_proxy_async();   // the previous document's generic smem accesses precede the async writes
        group_bar<GW>(grp);
        const int qi = __reduce_max_sync(STM_FULL, *qslot);
        if (qi >= P.n_docs) break;
        const int d = P.docs[qi];
        const long long p0 = P.doc_ptr[d];
        const int n = __reduce_max_sync(STM_FULL, (int)(P.doc_ptr[d + 1] - p0));
        const int asp = P.aspect ? P.aspect[d] : 0;
        const float* beta_a = P.beta_t + (size_t)asp * P.V * TS;
        double* beta_ss_a = P.beta_ss_t + (size_t)asp * P.V * TS;

        // ---- gather: ids / counts, TMA bulk copies of the beta rows -----------------------------
        if (gt == 0) mbar_expect_tx(mbar, (uint32_t)(n * TS * 4));
        double nsum_l = 0.0;
        for (int v = gt; v < n; v += POST_GT) {
            const int w = P.word_id[p0 + v];
            const float c = P.count[p0 + v];
            wid[v] = w * TS;   // row offset into beta_ss (V * TS < 2^31), so that the phi scatter needs no 64-bit multiply
            cw[v] = c;
            nsum_l += (double)c;
            tma_row_g2s(tile + (size_t)v * TS, beta_a + (size_t)w * TS, (uint32_t)(TS * 4), mbar);
        }
        if (gt < 8) tile[(size_t)n * TS + gt] = 0.f;   // what the last word's last topic block over-reads
        const double Nsum = group_sum<GW>(nsum_l, red, wg, lane, grp);   // np.sum(word_count)

        // ---- theta (stm.py:546-549, 905-909) and the prior quadratic, by the group's first warp ---
        double th[KPL];
        if (wg == 0) {
            double x[KPL], et[KPL], eu[KPL], m = -INFINITY, se_l = 0.0, ss_l = 0.0, q_l = 0.0;
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                x[i] = (k < K1) ? P.eta[(size_t)d * K1 + k] : 0.0;
                const double mu = (k < K1) ? P.mu[(size_t)d * K1 + k] : 0.0;
                const double dk = x[i] - mu;
                q_l += (Sd[i] * dk) * dk;
                et[i] = (k < K1) ? x[i] : ((k == K1) ? 0.0 : -INFINITY);
                m = nanmax(m, et[i]);
            }
            m = warp_max(m);
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                eu[i] = exp(et[i]);
                th[i] = exp(et[i] - m);
                se_l += eu[i];
                ss_l += th[i];
            }
            const double se = warp_sum(se_l), ss = warp_sum(ss_l);
            const double quad = warp_sum(q_l);
#pragma unroll
            for (int i = 0; i < KPL; ++i) {
                const int k = lane + 32 * i;
                th[i] = th[i] / ss;
                if (k < K) P.theta[(size_t)d * K + k] = eu[i] / se;  // stm.py:547-549 (no max shift)
                v0[k] = (k < K) ? eu[i] : 0.0;
                v1[k] = (k < K) ? th[i] * eu[i] : 0.0;
                v2[k] = (k < K) ? th[i] : 0.0;
            }
            if (lane == 0) { red[20] = quad; red[21] = 0.0; }   // red[21]: "the first PD test is already decided"
        }
        mbar_wait(mbar, parity);
        parity ^= 1;
        group_bar<GW>(grp);

        // ---- colsum_v = sum_k e_k beta_kv, log-likelihood part of the bound (stm.py:1088-1096) ----
        double loglik;
        {
            LogProd lp;
            logprod_init(lp);
            const double2* e2 = reinterpret_cast<const double2*>(v0);
            const double2* t2 = reinterpret_cast<const double2*>(v1);
#pragma unroll 1
            for (int v = gt; v < n; v += POST_GT) {
                const float4* row = reinterpret_cast<const float4*>(tile + (size_t)v * TS);
                double s0 = 0.0, s1 = 0.0, t0 = 0.0, t1 = 0.0;
#pragma unroll 1
                for (int q = 0; q < TS / 4; ++q) {
                    const float4 bq = row[q];
                    const double2 ea = e2[2 * q], eb = e2[2 * q + 1], wa = t2[2 * q], wb = t2[2 * q + 1];
                    const double b0 = beta_f2d(bq.x), b1 = beta_f2d(bq.y), b2 = beta_f2d(bq.z), b3 = beta_f2d(bq.w);
                    s0 = fma(ea.x, b0, s0); t0 = fma(wa.x, b0, t0);
                    s1 = fma(ea.y, b1, s1); t1 = fma(wa.y, b1, t1);
                    s0 = fma(eb.x, b2, s0); t0 = fma(wb.x, b2, t0);
                    s1 = fma(eb.y, b3, s1); t1 = fma(wb.y, b3, t1);
                }
                const float cf = cw[v];
                logprod_add(lp, t0 + t1, cf);
                logprod_renorm(lp);
                const double sq = sqrt((double)cf);
                wv[v] = sq / (s0 + s1);   // sqrt(c_v)/colsum_v
                wv2[v] = sq;
            }
            if (gt < 4) wv[n + gt] = 0.0;
            loglik = group_sum<GW>(logprod_value(lp), red, wg, lane, grp);   // also publishes wv / wv2
        }

        // ---- Hessian data term sum_v b_v b_v' (stm.py:1000-1006) on the fp64 tensor cores (DMMA
        // m8n8k4, four words per step), one 8-topic row block at a time; phi -> beta_ss
        // (stm.py:1103-1118, 582-590) and its row sums by the warp that owns the row block ---------
        {
            const int nbp = (K + 7) >> 3;
            const int w4 = lane & 3, kk = lane >> 2;
            double ek[NBMAX];   // exp(eta~) of this lane's topics (zero beyond K)
#pragma unroll
            for (int t = 0; t < NBMAX; ++t) ek[t] = v0[8 * t + kk];
            if constexpr (NBMAX > 8) {
                // the unit table (built once per CTA, below the kernel prologue): descending size, dealt in snake order
                const int nunits = unit_count_s;
                int slot = 0;
#pragma unroll 1
                for (int u = 0; u < nunits; ++u, ++slot) {
                    const int m6 = slot % (2 * POST_GW);
                    if ((m6 < POST_GW ? m6 : 2 * POST_GW - 1 - m6) != wg) continue;
                    const int code = __reduce_max_sync(STM_FULL, (int)units_s[u]);   // provably uniform
                    const int br = code & 15, c0 = 7 * ((code >> 4) & 3), nc = (code >> 8) & 7;
                    double acc[7][2], ek7[7];
#pragma unroll
                    for (int t = 0; t < 7; ++t) { acc[t][0] = 0.0; acc[t][1] = 0.0; ek7[t] = (t < nc) ? v0[8 * (c0 + t) + kk] : 0.0; }
                    double rs = 0.0;
                    const int kb = 8 * br + kk;
                    const double ekb = v0[kb];
                    const bool kok = kb < K;
                    const bool do_phi = (c0 == 0);
                    double* ssb = beta_ss_a + kb;
                    switch (nc) {
                        case 1: hess_unit_pass<1>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        case 2: hess_unit_pass<2>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        case 3: hess_unit_pass<3>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        case 4: hess_unit_pass<4>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        case 5: hess_unit_pass<5>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        case 6: hess_unit_pass<6>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                        default: hess_unit_pass<7>(tile, TS, n, wv, wv2, wid, ek7, ekb, kok, ssb, w4, kk, br, c0, do_phi, acc, rs); break;
                    }
                    const int gi = br * 8 + kk;
#pragma unroll
                    for (int t = 0; t < 7; ++t) {
                        if (t < nc) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int gj = (c0 + t) * 8 + 2 * w4 + c;
                                if (gi < K1 && gj <= gi) Hg[(size_t)gi * K1 + gj] = acc[t][c];
                            }
                        }
                    }
                    if (do_phi) {
                        rs += __shfl_xor_sync(STM_FULL, rs, 1);
                        rs += __shfl_xor_sync(STM_FULL, rs, 2);
                        if (w4 == 0 && kb < KV) v3[kb] = rs;
                    }
                }
            } else if (STM_HESS_FUSED && USE_TM && NBMAX == 8 && POST_GW == 3 && nbp == 7) {
                if constexpr (NBMAX == 8 && POST_GW == 3) {
                    double acc0[8][2], acc1[8][2], acc2[8][2];
#pragma unroll
                    for (int bc = 0; bc < 8; ++bc) {
                        acc0[bc][0] = 0.0; acc0[bc][1] = 0.0; acc1[bc][0] = 0.0; acc1[bc][1] = 0.0;
                        acc2[bc][0] = 0.0; acc2[bc][1] = 0.0;
                    }
                    double rs0 = 0.0, rs1 = 0.0, rs2 = 0.0;
                    int r0, r1, r2;      // the snake deal of block rows 6 .. 0 over three warps: {6,1,0} {5,2} {4,3}
                    if (wg == 0) {
                        hess_rows_fused<6, 1, 0>(tile, TS, n, wv, wv2, wid, ek, K, beta_ss_a, w4, kk, acc0, acc1, acc2, rs0, rs1, rs2);
                        r0 = 6; r1 = 1; r2 = 0;
                    } else if (wg == 1) {
                        hess_rows_fused<5, 2, -1>(tile, TS, n, wv, wv2, wid, ek, K, beta_ss_a, w4, kk, acc0, acc1, acc2, rs0, rs1, rs2);
                        r0 = 5; r1 = 2; r2 = -1;
                    } else {
                        hess_rows_fused<4, 3, -1>(tile, TS, n, wv, wv2, wid, ek, K, beta_ss_a, w4, kk, acc0, acc1, acc2, rs0, rs1, rs2);
                        r0 = 4; r1 = 3; r2 = -1;
                    }
                    // the same three tensor-memory slots, in the order the assembly below walks the warp's rows
#pragma unroll
                    for (int bc = 0; bc < 8; bc += 2) {
                        tm_st_d4(tmw + (uint32_t)(4 * bc), acc0[bc][0], acc0[bc][1], acc0[bc + 1][0], acc0[bc + 1][1]);
                        tm_st_d4(tmw + (uint32_t)(32 + 4 * bc), acc1[bc][0], acc1[bc][1], acc1[bc + 1][0], acc1[bc + 1][1]);
                        tm_st_d4(tmw + (uint32_t)(64 + 4 * bc), acc2[bc][0], acc2[bc][1], acc2[bc + 1][0], acc2[bc + 1][1]);
                    }
                    rs0 += __shfl_xor_sync(STM_FULL, rs0, 1); rs0 += __shfl_xor_sync(STM_FULL, rs0, 2);
                    rs1 += __shfl_xor_sync(STM_FULL, rs1, 1); rs1 += __shfl_xor_sync(STM_FULL, rs1, 2);
                    rs2 += __shfl_xor_sync(STM_FULL, rs2, 1); rs2 += __shfl_xor_sync(STM_FULL, rs2, 2);
                    if (w4 == 0) {
                        v3[8 * r0 + kk] = rs0;
                        if (r1 >= 0) v3[8 * r1 + kk] = rs1;
                        if (r2 >= 0) v3[8 * r2 + kk] = rs2;
                    }
                }
            } else {
            int qpos = 0, ps = 0;
#pragma unroll 1
            for (int br = nbp - 1; br >= 0; --br, ++qpos) {
                const int m6 = qpos % (2 * POST_GW);
                if ((m6 < POST_GW ? m6 : 2 * POST_GW - 1 - m6) != wg) continue;
                double acc[NBMAX][2];
#pragma unroll
                for (int bc = 0; bc < NBMAX; ++bc) { acc[bc][0] = 0.0; acc[bc][1] = 0.0; }
                double rs = 0.0;
                const int kb = 8 * br + kk;
                const double ekb = v0[kb];
                const bool kok = kb < K;
                double* ssb = beta_ss_a + kb;
                if constexpr (NBMAX > 8) {
                    HessDispatch<NBMAX, NBMAX - 1>::run(br, tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
                } else {
                    switch (br) {
                        case 0: hess_row_pass<NBMAX, 0>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs); break;
                        case 1: hess_row_pass<NBMAX, 1>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs); break;
                        case 2: hess_row_pass<NBMAX, 2>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs); break;
                        case 3: hess_row_pass<NBMAX, 3>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs); break;
                        default:
                            if constexpr (NBMAX > 4) {
                                if (br == 4) hess_row_pass<NBMAX, 4>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
                                else if (br == 5) hess_row_pass<NBMAX, 5>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
                                else if (br == 6) hess_row_pass<NBMAX, 6>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
                                else hess_row_pass<NBMAX, 7>(tile, TS, n, wv, wv2, wid, ek, ekb, kok, ssb, w4, kk, acc, rs);
                            }
                            break;
                    }
                }
                // C fragment: row = lane>>2, cols = 2*(lane&3) + {0,1}
                if constexpr (USE_TM) {
#pragma unroll
                    for (int bc = 0; bc < NBMAX; bc += 2)
                        tm_st_d4(tmw + (uint32_t)(ps * 4 * NBMAX + 4 * bc), acc[bc][0], acc[bc][1], acc[bc + 1][0], acc[bc + 1][1]);
                    ++ps;
                } else {
                    const int gi = br * 8 + kk;
#pragma unroll
                    for (int bc = 0; bc < NBMAX; ++bc) {
                        if (bc <= br) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int gj = bc * 8 + 2 * w4 + c;
                                if (gi < K1 && gj <= gi) Hg[(size_t)gi * K1 + gj] = acc[bc][c];
                            }
                        }
                    }
                }
                rs += __shfl_xor_sync(STM_FULL, rs, 1);
                rs += __shfl_xor_sync(STM_FULL, rs, 2);
                if (w4 == 0 && kb < KV) v3[kb] = rs;
            }
        }
            }
        if constexpr (USE_TM) tm_wait_st();
        group_bar<GW>(grp);   // tile dead; Hg and v3 complete

        // ---- assemble H = data - N theta theta' + diag(-rowsum + N theta) + siginv (stm.py:1007-1015)
        // (the data term comes back from the L2 bounce.  The threads walk the packed lower triangle, element gt, gt + GT,
        // ...: (row, column) is stepped forward by GT elements per visit — no square root per element — and the loads
        // of STM_ASM_BATCH visits are issued together so that one L2 round trip covers them)
        if constexpr (USE_TM) {
            // every warp takes its own row passes back out of tensor memory and assembles their elements
            const int nbp = (K + 7) >> 3;
            const int w4 = lane & 3, kk = lane >> 2;
            int qpos = 0, ps = 0;
#pragma unroll 1
            for (int br = nbp - 1; br >= 0; --br, ++qpos) {
                const int m6 = qpos % (2 * POST_GW);
                if ((m6 < POST_GW ? m6 : 2 * POST_GW - 1 - m6) != wg) continue;
                uint32_t raw[NBMAX / 2][8];
#pragma unroll
                for (int b2 = 0; b2 < NBMAX / 2; ++b2) tm_ld8(tmw + (uint32_t)(ps * 4 * NBMAX + 8 * b2), raw[b2]);
                tm_wait_ld();
                ++ps;
                double acc[NBMAX][2];
#pragma unroll
                for (int b2 = 0; b2 < NBMAX / 2; ++b2)
                    tm_d4_from(raw[b2], acc[2 * b2][0], acc[2 * b2][1], acc[2 * b2 + 1][0], acc[2 * b2 + 1][1]);
                const int gi = br * 8 + kk;
                if (gi < K1) {
                    const double thr = v2[gi];
#pragma unroll
                    for (int bc = 0; bc < NBMAX; ++bc) {
                        if (bc <= br) {
#pragma unroll
                            for (int c = 0; c < 2; ++c) {
                                const int gj = bc * 8 + 2 * w4 + c;
                                if (gj <= gi) {
                                    const double thk = v2[gj];
                                    double h = acc[bc][c] - Nsum * (thr * thk);
                                    if (gj == gi) {
                                        h = (h - v3[gj] + Nsum * thk) + P.prior[gj];
                                        Dg[gj] = h;
                                        if (!(h > 0.0)) red[21] = 1.0;
                                    }
                                    Hm[(size_t)gi * HS + gj] = h;
                                    Hm[(size_t)gj * HS + gi] = h;
                                }
                            }
                        }
                    }
                }
            }
        } else {
            const int ntri = K1 * (K1 + 1) / 2;
            int r = 0, k = gt;                       // element gt of the packed triangle = (r, k)
            while (k > r) { k -= r + 1; ++r; }
            for (int base_i = gt; base_i < ntri; base_i += POST_GT * STM_ASM_BATCH) {
                double hv[STM_ASM_BATCH];
                int rr[STM_ASM_BATCH], kk_[STM_ASM_BATCH];
#pragma unroll
                for (int u = 0; u < STM_ASM_BATCH; ++u) {
                    rr[u] = r; kk_[u] = k;
                    hv[u] = (base_i + u * POST_GT < ntri) ? __ldcg(&Hg[(size_t)r * K1 + k]) : 0.0;
                    k += POST_GT;
                    while (k > r) { k -= r + 1; ++r; }
                }
#pragma unroll
                for (int u = 0; u < STM_ASM_BATCH; ++u) {
                    if (base_i + u * POST_GT < ntri) {
                        const int r_ = rr[u], k_ = kk_[u];
                        const double thk = v2[k_];
                        double h = hv[u] - Nsum * (v2[r_] * thk);
                        if (k_ == r_) {
                            h = (h - v3[k_] + Nsum * thk) + P.prior[k_];
                            Dg[k_] = h;
                            if (!(h > 0.0)) red[21] = 1.0;
                        }
                        Hm[(size_t)r_ * HS + k_] = h;
                        Hm[(size_t)k_ * HS + r_] = h;
                    }
                }
            }
        }
        group_bar<GW>(grp);
        // A 2x2 principal minor that is negative by a clear margin (H_ij^2 > H_ii H_jj) also decides the first
        // PD test without a sweep: in the spectral-init state ~90 % of the documents fail that test, on average
        // two thirds of the way through the pivots (measured on the oracle).  Margin 1e-10: anything closer
        // is left to the pivots.
        if (red[21] == 0.0) {
            bool neg = false;
            for (int r = wg; r < K1; r += POST_GW) {
                const double dr = Dg[r] * (1.0 + 1e-10);
                for (int k = lane; k < r; k += 32) {
                    const double h = Hm[(size_t)r * HS + k];
                    neg |= (h * h > dr * Dg[k]);
                }
            }
            if (neg) red[21] = 1.0;
        }
        group_bar<GW>(grp);

        // ---- PD test + repairs (stm.py:1017-1021, 1039-1048) and the inverse, by sweeping ----------
        int repair = 0, upper = 0, dead = 0;
        double a[4][4];
        for (int attempt = 0;; ++attempt) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * pi + r, j = 4 * pj + c;
                    a[r][c] = (has_patch && i < K1 && j < K1) ? ((i == j) ? Dg[i] : Hm[(size_t)i * HS + j])
                                                              : ((i == j) ? 1.0 : 0.0);
                }
            // a non-positive diagonal entry already decides the first PD test (stm.py:1017): no sweep needed
            int ok = (attempt == 0 && red[21] != 0.0) ? 0 : 1;
#pragma unroll 1
            for (int kb = 0; 4 * kb < K1 && ok; ++kb) {
#pragma unroll
                for (int kr = 0; kr < 4; ++kr) {
                    const int k = 4 * kb + kr;
                    if (k < K1 && ok) {
                        double* ub = ubuf + (k & 1) * POST_UST;
                        if (has_patch) {
                            if (pi == kb) {
#pragma unroll
                                for (int c = 0; c < 4; ++c) ub[4 * pj + c] = a[kr][c];
                            } else if (pj == kb) {
#pragma unroll
                                for (int r = 0; r < 4; ++r) ub[4 * pi + r] = a[r][kr];
                            }
                        }
                        group_bar<GW>(grp);
                        const double dk = ub[k];
                        if (!(dk > 0.0)) {
                            ok = 0;
                        } else {
                            if (gt == 0) dvec[k] = dk;
                            const double pv = 1.0 / dk;   // inlined: on the sweep's critical path (r01 A/B: 10.4 -> 9.7 ms)
                            if (has_patch) {
                                double ui[4], upj[4];
#pragma unroll
                                for (int r = 0; r < 4; ++r) { ui[r] = ub[4 * pi + r]; upj[r] = ub[4 * pj + r] * pv; }
#pragma unroll
                                for (int r = 0; r < 4; ++r)
#pragma unroll
                                    for (int c = 0; c < 4; ++c) a[r][c] = fma(-ui[r], upj[c], a[r][c]);
                                if (pi == kb) {
#pragma unroll
                                    for (int c = 0; c < 4; ++c) a[kr][c] = upj[c];
                                }
                                if (pj == kb) {
#pragma unroll
                                    for (int r = 0; r < 4; ++r) a[r][kr] = ui[r] * pv;
                                    if (pi == kb) a[kr][kr] = -pv;
                                }
                            }
                        }
                    }
                }
            }
            if (ok) break;
            group_bar<GW>(grp);   // every thread has left the sweep before Dg changes
            if (attempt == 0 || attempt == 2 || attempt == 3) {
                // make_pd (stm.py:964-984): d_i <- max(d_i, sum_j!=i |H_ij|)
                for (int i = gt; i < K1; i += POST_GT) {
                    double tot = 0.0;
                    for (int k = 0; k < K1; ++k) if (k != i) tot += fabs(Hm[(size_t)i * HS + k]);
                    tot += fabs(Dg[i]);
                    const double mag = tot - fabs(Dg[i]);
                    if (Dg[i] < mag) Dg[i] = mag;
                }
            }
            if (attempt == 1 || attempt == 3) {
                for (int i = gt; i < K1; i += POST_GT) Dg[i] += 1e-5;
            }
            if (attempt == 0) repair = 1;            // hessian(): make_pd
            else if (attempt == 1) repair = 2;       // hessian(): + 1e-5
            else if (attempt == 2) repair += 4;      // decompose_hessian(): make_pd
            else if (attempt == 3) { repair += 8; upper = 1; }  // scipy.linalg.cholesky (upper) of make_pd + 1e-5 I
            else { dead = 1; break; }
            group_bar<GW>(grp);
        }
        group_bar<GW>(grp);   // dvec complete

        // ---- bound (stm.py:1085-1100): sum log L_ii = 1/2 sum log d_k ------------------------------
        double lg = 0.0;
        if (gt < K1) lg = dead ? nan("") : log_noinline(dvec[gt]);
        const double logdet_half = 0.5 * group_sum<GW>(lg, red, wg, lane, grp);
        if (gt == 0) {
            P.doc_bound[d] = loglik - logdet_half - 0.5 * red[20] - sigmaentropy;
            P.doc_info[d] = (P.doc_info[d] & 0xffffff) | (repair << 24);
        }
        // ---- nu = H^-1 (stm.py:1052-1066, accumulated stm.py:582): the swept matrix is -H^-1 --------
        if (has_patch) {
#pragma unroll
            for (int r = 0; r < 4; ++r)
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int i = 4 * pi + r, j = 4 * pj + c;
                    if (i < K1 && j <= i) {
                        double nu;
                        if (dead) nu = nan("");
                        else if (!upper) nu = -a[r][c];
                        // scipy's UPPER factor in optimize_nu: only its diagonal survives np.triu(L.T)
                        else nu = (i == j) ? 1.0 / dvec[i] : 0.0;
                        if (!upper || i == j) red_add_f64(sig_acc + (size_t)i * K1 + j, nu);
                    }
                }
        }
    }  // document loop
    if constexpr (USE_TM) {
        asm volatile("tcgen05.fence::before_thread_sync;");
        __syncthreads();
        if (warp_u == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tm_base_b));
    }
}

}  // namespace stm
