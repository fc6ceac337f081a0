// corpus_gen.cuh — synthetic corpus sampler on the device.
// Replaces CorpusCreation.sample_documents, /root/reference/src/modules/generate_docs.py:293-316 (dgp = "STM"
// / "LDA": document d ~ Multinomial(n_words, theta_d beta)), without materialising the D x V matrix
// theta @ beta (generate_docs.py:297 — 8 GB at D=100k, V=10k): each token draws its topic from theta_d and
// then its word from beta_z, which is the same multinomial.  Included by stm_b200.cu.
//
// Random numbers: Philox4x32-10 (Salmon, Moraes, Dror, Shaw: "Parallel random numbers: as easy as 1, 2, 3",
// SC'11), counter = (token, doc_lo, doc_hi, 0), key = seed — stateless, so the CPU oracle
// (oracle/corpus_numpy.py) reproduces the corpus bit for bit.  The reference's own stream (NumPy's PCG64
// multinomial) cannot be matched by any parallel sampler; parity for this row is the exact restatement of
// THIS sampler plus distributional tests against theta @ beta.
#pragma once

namespace stm_gen {

__device__ __forceinline__ uint4 philox4x32_10(uint4 c, uint2 k) {
#pragma unroll
    for (int r = 0; r < 10; ++r) {
        const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
        const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
        c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
        k.x += 0x9E3779B9u;
        k.y += 0xBB67AE85u;
    }
    return c;
}
// 53-bit uniform in [0, 1)
__device__ __forceinline__ double u53(uint32_t hi, uint32_t lo) {
    return (double)((((unsigned long long)hi << 32) | lo) >> 11) * 1.1102230246251565e-16;
}

// np.cumsum along each row of beta (sequential order: bit-identical to NumPy); one thread per topic
__global__ void cum_rows_kernel(const double* __restrict__ beta, int K, int V, double* __restrict__ cum) {
    const int k = blockIdx.x * blockDim.x + threadIdx.x;
    if (k >= K) return;
    double s = 0.0;
    for (int v = 0; v < V; ++v) {
        s += beta[(size_t)k * V + v];
        cum[(size_t)k * V + v] = s;
    }
}

// first index with cum[i] > x (np.searchsorted(cum, x, side="right")), clamped to n-1
__device__ __forceinline__ int upper_bound(const double* __restrict__ cum, int n, double x) {
    int lo = 0, hi = n;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cum[mid] > x) hi = mid; else lo = mid + 1;
    }
    return lo < n ? lo : n - 1;
}

// One CTA per document: n_words (topic, word) draws, bitonic sort of the word ids in shared memory,
// number of distinct words -> n_unique[d]; the sorted tokens go to tok[d][0..n_words).
__global__ void sample_tokens_kernel(const double* __restrict__ theta, const double* __restrict__ cum_beta, int K,
                                     int V, int n_words, int npow2, unsigned long long seed, long long D,
                                     int* __restrict__ tok, int* __restrict__ n_unique) {
    extern __shared__ int sh[];            // npow2 ints, then K doubles (8-byte aligned: npow2 is even)
    double* cth = reinterpret_cast<double*>(sh + npow2);
    __shared__ int heads;
    const long long d = blockIdx.x;
    if (d >= D) return;
    if (threadIdx.x == 0) {
        double s = 0.0;
        for (int k = 0; k < K; ++k) { s += theta[(size_t)d * K + k]; cth[k] = s; }
        heads = 0;
    }
    __syncthreads();
    const double ttot = cth[K - 1];
    const uint2 key = make_uint2((uint32_t)seed, (uint32_t)(seed >> 32));
    for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
        int w = 0x7fffffff;                // padding sorts to the end
        if (t < n_words) {
            const uint4 r = philox4x32_10(make_uint4((uint32_t)t, (uint32_t)d, (uint32_t)((unsigned long long)d >> 32), 0u), key);
            const int z = upper_bound(cth, K, u53(r.x, r.y) * ttot);
            const double* cb = cum_beta + (size_t)z * V;
            w = upper_bound(cb, V, u53(r.z, r.w) * cb[V - 1]);
        }
        sh[t] = w;
    }
    __syncthreads();
    for (int k = 2; k <= npow2; k <<= 1)
        for (int j = k >> 1; j > 0; j >>= 1) {
            for (int t = threadIdx.x; t < npow2; t += blockDim.x) {
                const int p = t ^ j;
                if (p > t) {
                    const int a = sh[t], b = sh[p];
                    const bool up = ((t & k) == 0);
                    if ((a > b) == up) { sh[t] = b; sh[p] = a; }
                }
            }
            __syncthreads();
        }
    int mine = 0;
    for (int t = threadIdx.x; t < n_words; t += blockDim.x) {
        tok[(size_t)d * n_words + t] = sh[t];
        mine += (t == 0 || sh[t] != sh[t - 1]);
    }
    atomicAdd(&heads, mine);
    __syncthreads();
    if (threadIdx.x == 0) n_unique[d] = heads;
}

// doc_ptr = exclusive prefix sum of n_unique (one CTA, contiguous chunk per thread)
__global__ void scan_kernel(const int* __restrict__ n_unique, long long D, long long* __restrict__ doc_ptr) {
    __shared__ long long part[1024];
    const long long chunk = (D + blockDim.x - 1) / blockDim.x;
    const long long lo = (long long)threadIdx.x * chunk, hi = min(D, lo + chunk);
    long long s = 0;
    for (long long i = lo; i < hi; ++i) s += n_unique[i];
    part[threadIdx.x] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        long long run = 0;
        for (int i = 0; i < (int)blockDim.x; ++i) { const long long v = part[i]; part[i] = run; run += v; }
        doc_ptr[D] = run;
    }
    __syncthreads();
    long long run = part[threadIdx.x];
    for (long long i = lo; i < hi; ++i) { doc_ptr[i] = run; run += n_unique[i]; }
}

// run-length encode the sorted tokens of each document into the CSR arrays (one warp per document)
__global__ void rle_kernel(const int* __restrict__ tok, int n_words, long long D, const long long* __restrict__ doc_ptr,
                           int* __restrict__ word_id, float* __restrict__ count) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    for (long long d = warp; d < D; d += nwarps) {
        const int* t = tok + (size_t)d * n_words;
        long long base = doc_ptr[d];
        for (int t0 = 0; t0 < n_words; t0 += 32) {
            const int i = t0 + lane;
            const bool head = i < n_words && (i == 0 || t[i] != t[i - 1]);
            const unsigned m = __ballot_sync(0xffffffffu, head);
            if (head) {
                int e = i + 1;
                while (e < n_words && t[e] == t[i]) ++e;
                const long long pos = base + __popc(m & ((1u << lane) - 1u));
                word_id[pos] = t[i];
                count[pos] = (float)(e - i);
            }
            base += __popc(m);
        }
    }
}

}  // namespace stm_gen

extern "C" {

int stm_sample_corpus(stm_ctx* ctx, int64_t D, int n_words, const double* theta_dev, const double* beta_kv_dev,
                      uint64_t seed, int64_t* doc_ptr_dev, int32_t* word_id_dev, float* count_dev, int64_t* nnz_out,
                      void* stream) {
    using namespace stm_gen;
    if (!ctx) return STM_ERR_INVALID;
    if (D < 1 || n_words < 1 || n_words > 4096 || !theta_dev || !beta_kv_dev || !doc_ptr_dev || !word_id_dev ||
        !count_dev || !nnz_out)
        return fail(ctx, STM_ERR_INVALID, "stm_sample_corpus: bad arguments (1 <= n_words <= 4096)");
    if (D > 0x7fffffffLL) return fail(ctx, STM_ERR_UNSUPPORTED, "stm_sample_corpus: more than 2^31 documents");
    STM_ON_DEVICE(ctx);
    cudaStream_t st = (cudaStream_t)stream;
    const int K = ctx->K, V = ctx->V;
    int npow2 = 2;
    while (npow2 < n_words) npow2 <<= 1;
    double* cum = nullptr; int *tok = nullptr, *nuniq = nullptr;
    auto cleanup = [&]() { cudaFree(cum); cudaFree(tok); cudaFree(nuniq); };
#define GCU(call)                                                                                    \
    do {                                                                                             \
        cudaError_t e_ = (call);                                                                     \
        if (e_ != cudaSuccess) {                                                                     \
            cleanup();                                                                               \
            return fail(ctx, STM_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_));      \
        }                                                                                            \
    } while (0)
    GCU(cudaMalloc(&cum, sizeof(double) * (size_t)K * V));
    GCU(cudaMalloc(&tok, sizeof(int) * (size_t)D * n_words));
    GCU(cudaMalloc(&nuniq, sizeof(int) * (size_t)D));
    cum_rows_kernel<<<(K + 31) / 32, 32, 0, st>>>(beta_kv_dev, K, V, cum);
    const size_t smem = sizeof(int) * npow2 + sizeof(double) * K;
    sample_tokens_kernel<<<(unsigned)D, 128, smem, st>>>(theta_dev, cum, K, V, n_words, npow2, seed, D, tok, nuniq);
    scan_kernel<<<1, 1024, 0, st>>>(nuniq, D, reinterpret_cast<long long*>(doc_ptr_dev));
    rle_kernel<<<ctx->sm_count * 8, 256, 0, st>>>(tok, n_words, D, reinterpret_cast<const long long*>(doc_ptr_dev),
                                                  word_id_dev, count_dev);
    ctx->launches += 4;
    long long nnz = 0;
    GCU(cudaMemcpyAsync(&nnz, doc_ptr_dev + D, sizeof(long long), cudaMemcpyDeviceToHost, st));
    GCU(cudaStreamSynchronize(st));
    GCU(cudaGetLastError());
#undef GCU
    cleanup();
    *nnz_out = nnz;
    return STM_OK;
}

}  // extern "C"
