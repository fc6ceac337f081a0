// estep_inst.cu — instantiates stm::bfgs_kernel<STM_KPL, J> for J in {2,4,5,8} and
// stm::post_kernel<STM_KPL>; compiled once per STM_KPL in {1,2,3,4} (K <= 32*STM_KPL) so the
// instantiations build in parallel.
#include "estep_kernel.cuh"

#ifndef STM_KPL
#error "compile with -DSTM_KPL=1..4"
#endif
#define STM_CAT2(a, b) a##b
#define STM_CAT(a, b) STM_CAT2(a, b)

cudaError_t STM_CAT(stm_launch_bfgs_kpl, STM_KPL)(const stm::EstepParams& P, int J, int grid, int block,
                                                  size_t smem, cudaStream_t st) {
#define STM_LAUNCH(JJ)                                                                                \
    {                                                                                                 \
        cudaError_t e = cudaFuncSetAttribute(stm::bfgs_kernel<STM_KPL, JJ>,                           \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
        if (e != cudaSuccess) return e;                                                               \
        stm::bfgs_kernel<STM_KPL, JJ><<<grid, block, smem, st>>>(P);                                  \
        return cudaGetLastError();                                                                    \
    }
    switch (J) {
        case 2: STM_LAUNCH(2)
        case 4: STM_LAUNCH(4)
        case 5: STM_LAUNCH(5)
        default: STM_LAUNCH(8)
    }
#undef STM_LAUNCH
}

cudaError_t STM_CAT(stm_launch_post_kpl, STM_KPL)(const stm::EstepParams& P, int grid, int block, size_t smem,
                                                  cudaStream_t st) {
    cudaError_t e = cudaFuncSetAttribute(stm::post_kernel<STM_KPL>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem);
    if (e != cudaSuccess) return e;
    stm::post_kernel<STM_KPL><<<grid, block, smem, st>>>(P);
    return cudaGetLastError();
}

// group version of kernel B (GW warps per document, one 4x4 patch of the lower triangle per thread):
// the (KPL, GW) pairs that occur — K-1 <= 52: GW 3; K-1 <= 63: 5; K <= 96: 10; K-1 <= 100: 11; K <= 128: 17
cudaError_t STM_CAT(stm_launch_post_group_kpl, STM_KPL)(const stm::EstepParams& P, int gw, int grid, int block,
                                                        size_t smem, cudaStream_t st) {
#define STM_LAUNCH_G(GW)                                                                               \
    if (gw == GW) {                                                                                    \
        cudaError_t e = cudaFuncSetAttribute(stm::post_group_kernel<STM_KPL, GW>,                      \
                                             cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);  \
        if (e != cudaSuccess) return e;                                                                \
        stm::post_group_kernel<STM_KPL, GW><<<grid, block, smem, st>>>(P);                             \
        return cudaGetLastError();                                                                     \
    }
#if STM_KPL <= 2
    STM_LAUNCH_G(3)
#endif
#if STM_KPL == 2
    STM_LAUNCH_G(5)
#endif
#if STM_KPL == 3
    STM_LAUNCH_G(10)
#endif
#if STM_KPL == 4
    STM_LAUNCH_G(11)
    STM_LAUNCH_G(17)
#endif
#undef STM_LAUNCH_G
    (void)P; (void)grid; (void)block; (void)smem; (void)st;
    return cudaErrorNotSupported;
}
