// lat_probe.cu — dependent-chain latencies (cycles) of the fp64 building blocks of kernel A on B200:
// DFMA, DADD, fp64 division / sqrt / log / exp (out of line, as the kernel calls them), SHFL, LDS.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o lat_probe lat_probe.cu && ./lat_probe
#include <cstdio>
#include <cuda_runtime.h>
static __device__ __noinline__ double ddiv(double a, double b) { return a / b; }
static __device__ __noinline__ double dsqrt_(double a) { return sqrt(a); }
static __device__ __noinline__ double dlog(double a) { return log(a); }
static __device__ __noinline__ double dexp(double a) { return exp(a); }
template <int OP>
__global__ void probe(double* out, long long* cyc, double seed, int n) {
    __shared__ double sh[64];
    sh[threadIdx.x & 63] = seed;
    __syncthreads();
    double x = seed + threadIdx.x * 1e-9, y = 1.0000001;
    long long t0 = clock64();
    for (int i = 0; i < n; ++i) {
        if (OP == 0) x = fma(x, y, 1e-9);
        else if (OP == 1) x = x + y;
        else if (OP == 2) x = ddiv(x, y);
        else if (OP == 3) x = x / y;
        else if (OP == 4) x = dsqrt_(x) + 1.0;
        else if (OP == 5) x = dlog(x) + 3.0;
        else if (OP == 6) x = dexp(x) * 0.3;
        else if (OP == 7) x = __shfl_xor_sync(0xffffffffu, x, 1);
        else if (OP == 8) { x = sh[((int)x) & 63]; }
        else if (OP == 9) x = x * y;
        else if (OP == 10) { x = (x > y) ? x - 1e-9 : x + 1e-9; }
        else if (OP == 11) x = log(x) + 3.0;
        else if (OP == 12) x = exp(x) * 0.3;
    }
    long long t1 = clock64();
    out[threadIdx.x + blockIdx.x * blockDim.x] = x;
    if (threadIdx.x == 0) cyc[blockIdx.x] = t1 - t0;
}
template <int OP>
void run(const char* name, int warps) {
    double* out; long long* cyc;
    cudaMalloc(&out, 8 * 1024 * 148); cudaMalloc(&cyc, 8 * 148);
    const int n = 2000;
    probe<OP><<<1, 32 * warps>>>(out, cyc, 1.5, n);
    probe<OP><<<1, 32 * warps>>>(out, cyc, 1.5, n);
    long long c; cudaMemcpy(&c, cyc, 8, cudaMemcpyDeviceToHost);
    printf("%-28s warps/CTA %2d : %7.1f cycles per op (dependent chain)\n", name, warps, (double)c / n);
    cudaFree(out); cudaFree(cyc);
}
int main() {
    for (int w : {1, 4, 8, 16}) {
        run<0>("DFMA", w); run<1>("DADD", w); run<9>("DMUL", w); run<10>("DSETP+select", w);
        run<2>("fp64 div (noinline call)", w); run<3>("fp64 div (inline)", w); run<4>("sqrt (noinline) + add", w);
        run<5>("log (noinline) + add", w); run<11>("log (inline) + add", w); run<6>("exp (noinline) * c", w); run<12>("exp (inline) * c", w);
        run<7>("SHFL.BFLY f64 (2x32)", w); run<8>("LDS f64 dependent", w);
    }
    return 0;
}
