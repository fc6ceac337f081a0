// TMEM probe (development tool): checks the lane/column addressing of tcgen05.st/ld .32x32b from
// several warps and measures tcgen05.ld throughput.  nvcc -gencode arch=compute_100a,code=sm_100a
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

__device__ __forceinline__ void tm_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tm_ld8(uint32_t taddr, uint32_t (&v)[8]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile("tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
                 : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]),
                   "=r"(v[8]), "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
                 : "r"(taddr)
                 : "memory");
}
__device__ __forceinline__ void tm_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tm_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

__global__ void probe(int active_warps, int iters, int* errs, long long* cyc, uint32_t* sink) {
    __shared__ uint32_t tm_base;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (warp == 0) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(
            (uint32_t)__cvta_generic_to_shared(&tm_base)));
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;");
    }
    asm volatile("tcgen05.fence::before_thread_sync;");
    __syncthreads();
    asm volatile("tcgen05.fence::after_thread_sync;");
    const uint32_t base = tm_base;
    if (threadIdx.x == 0) printf("tmem base = 0x%08x\n", base);
    const int q = warp & 3, half = warp >> 2;          // lane quarter, column half
    const uint32_t my = base + ((uint32_t)(q * 32) << 16) + (uint32_t)(half * 256);
    // write 256 columns: value encodes (warp, lane, col)
    for (int c = 0; c < 256; c += 8) {
        uint32_t v[8];
        for (int i = 0; i < 8; ++i) v[i] = (warp << 24) | (lane << 16) | (c + i);
        tm_st8(my + c, v);
    }
    tm_wait_st();
    __syncthreads();
    int bad = 0;
    for (int c = 0; c < 256; c += 8) {
        uint32_t v[8];
        tm_ld8(my + c, v);
        tm_wait_ld();
        for (int i = 0; i < 8; ++i) bad += (v[i] != ((warp << 24) | (lane << 16) | (c + i)));
    }
    atomicAdd(errs, bad);
    __syncthreads();
    // throughput: each active warp streams its 256 columns `iters` times
    uint32_t acc = 0;
    long long t0 = clock64();
    if ((active_warps >> warp) & 1) {
        for (int it = 0; it < iters; ++it) {
            uint32_t a[16], b[16];
            tm_ld16(my, a);
            for (int c = 16; c < 256; c += 32) {
                tm_wait_ld();
                tm_ld16(my + c, b);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += a[i];
                tm_wait_ld();
                if (c + 16 < 256) tm_ld16(my + c + 16, a);
#pragma unroll
                for (int i = 0; i < 16; ++i) acc += b[i];
            }
        }
    }
    long long t1 = clock64();
    if (lane == 0) cyc[warp] = t1 - t0;
    sink[threadIdx.x] = acc;
    __syncthreads();
    if (warp == 0) asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(base));
}

int main() {
    int* errs; long long* cyc; uint32_t* sink;
    cudaMalloc(&errs, 4); cudaMalloc(&cyc, 8 * 8); cudaMalloc(&sink, 4 * 256);
    cudaMemset(errs, 0, 4);
    for (int aw : {0x1, 0x11, 0x3, 0xF, 0xFF}) {
        probe<<<1, 256>>>(aw, 1000, errs, cyc, sink);
        cudaError_t e = cudaDeviceSynchronize();
        int h_err; long long h_cyc[8];
        cudaMemcpy(&h_err, errs, 4, cudaMemcpyDeviceToHost);
        cudaMemcpy(h_cyc, cyc, 64, cudaMemcpyDeviceToHost);
        const double bytes = 1000.0 * 256 * 32 * 4;   // per warp
        printf("active mask 0x%x: %s errs %d  cycles warp0 %lld  -> %.1f B/cyc/warp, %.1f B/cyc total\n", aw,
               cudaGetErrorString(e), h_err, h_cyc[0], bytes / h_cyc[0], __builtin_popcount(aw) * bytes / h_cyc[0]);
    }
    return 0;
}
