// I-cache scatter probe: the same HOT code (NCH chunks of H instructions) laid out contiguously, or with a cold block
// of C instructions (guarded by a run-time-false flag) after every chunk, so that the hot lines are scattered over a
// long function the way kernel A's hot step is scattered between its once-per-document code.  Run under
//   ncu --metrics sm__icc_request_hit_rate.pct,sm__icc_requests.sum,gcc__cache_requests_type_instruction.sum,gpu__time_duration.sum
#include <cstdio>
#include <cuda_runtime.h>

template <int N>
__device__ __forceinline__ void burn(float& a, float b, float c) {
#pragma unroll
    for (int i = 0; i < N; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
}

template <int H, int C, int NCH>
__global__ void __launch_bounds__(384, 1) scatter_kernel(int iters, const int* flags, float* out) {
    float a = threadIdx.x * 1e-3f, b = 1.0001f, c = 1e-7f;
    const int warp = threadIdx.x >> 5;
    for (int i = 0; i < warp * 97; ++i) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a) : "f"(b), "f"(c));
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < NCH; ++ch) {
            burn<H>(a, b, c);
            if (C > 0 && flags[ch] != 0) burn<(C > 0 ? C : 1)>(a, b, c);
        }
    }
    out[blockIdx.x * blockDim.x + threadIdx.x] = a;
}

template <int H, int C, int NCH>
void run(int warps, const int* flags, float* out) {
    const int iters = (1 << 21) / (H * NCH);
    scatter_kernel<H, C, NCH><<<148, warps * 32>>>(2, flags, out);
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    cudaEventRecord(e0);
    scatter_kernel<H, C, NCH><<<148, warps * 32>>>(iters, flags, out);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms; cudaEventElapsedTime(&ms, e0, e1);
    printf("hot %2d x %4d instr (%2d KB)  cold gap %5d instr (%3d KB span)  warps %2d : %.3f ms\n", NCH, H, NCH * H * 16 / 1024,
           C, NCH * (H + C) * 16 / 1024, warps, ms);
}

int main() {
    int* flags; float* out;
    cudaMalloc(&flags, 4 * 64); cudaMemset(flags, 0, 4 * 64);
    cudaMalloc(&out, 4 * 384 * 148);
    for (int warps : {1, 12}) {
        run<64, 0, 20>(warps, flags, out);      // 20 KB hot, contiguous
        run<64, 64, 20>(warps, flags, out);     // 20 KB hot over 40 KB
        run<64, 192, 20>(warps, flags, out);    // 20 KB hot over 80 KB
        run<64, 320, 20>(warps, flags, out);    // 20 KB hot over 120 KB
        run<32, 160, 40>(warps, flags, out);    // 20 KB hot in 0.5 KB pieces over 120 KB
        run<64, 0, 12>(warps, flags, out);      // 12 KB hot, contiguous
        run<64, 320, 12>(warps, flags, out);    // 12 KB hot over 72 KB
        run<64, 0, 28>(warps, flags, out);      // 28 KB hot, contiguous
        run<64, 192, 28>(warps, flags, out);    // 28 KB hot over 112 KB
    }
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
