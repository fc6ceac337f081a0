// I-cache probe (development tool): W warps per CTA loop over a straight-line body of N fp32 FMAs
// (16 B each) at staggered phases; reports cycles per warp-instruction vs body size, for 1 CTA and
// for one CTA per SM (shared instruction-cache levels).
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__global__ void __launch_bounds__(448, 1) body_kernel(int iters, float* out, long long* cyc) {
    const int warp = threadIdx.x >> 5;
    float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 1e-7f;
    // stagger the phases: warp w first burns w/W of a body's worth of time in a tiny loop
    for (int i = 0; i < warp * (N / (int)(blockDim.x >> 5)); ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    }
    const long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) cyc[blockIdx.x * (blockDim.x >> 5) + warp] = t1 - t0;
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <int N>
void run(int warps, int grid, float* out, long long* cyc) {
    const int iters = (1 << 22) / N;   // ~4M instructions per warp
    body_kernel<N><<<grid, warps * 32>>>(4, out, cyc);   // warm
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    body_kernel<N><<<grid, warps * 32>>>(iters, out, cyc);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    const double instr = (double)iters * N;            // per warp
    printf("body %4d KB  warps %2d grid %3d : %.3f ms, %.2f cycles/warp-instr (per warp), SM IPC %.2f\n", N * 16 / 1024, warps,
           grid, ms, ms * 1e-3 * 1.965e9 / instr, instr * warps / (ms * 1e-3 * 1.965e9));
}

int main() {
    float* out; long long* cyc;
    cudaMalloc(&out, 4 * 448 * 148); cudaMalloc(&cyc, 8 * 14 * 148);
    for (int grid : {1, 148}) {
        for (int warps : {4, 14}) {
            run<512>(warps, grid, out, cyc);
            run<1024>(warps, grid, out, cyc);
            run<1536>(warps, grid, out, cyc);
            run<2048>(warps, grid, out, cyc);
            run<3072>(warps, grid, out, cyc);
            run<4096>(warps, grid, out, cyc);
            run<6144>(warps, grid, out, cyc);
            run<8192>(warps, grid, out, cyc);
        }
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
