// The plain K-step kernels of temporal blocking (stepK_lean, stepK_stream) for ONE dtype and a range of K:
// compiled several times with -DFDTD_PART_T=double|float -DFDTD_K_LO=.. -DFDTD_K_HI=..  See launch.h.
#include "launch.h"

#ifndef FDTD_PART_T
#error "compile with -DFDTD_PART_T=double|float -DFDTD_K_LO=a -DFDTD_K_HI=b"
#endif

namespace fdtd {

template <typename T, int K>
cudaError_t launch_stepK_k(const StepParams<T>& P, int variant, int warps, dim3 grid, const unsigned char* skip, int ntx2,
                           cudaStream_t st) {
    constexpr int V2 = sizeof(T) == 8 ? 2 : 4;
    if (variant == 3) {                                          // ... with the stage chain cut (kernels.cuh stepK_skew_umu)
        const size_t lean_bytes = (size_t)warps * 4 * 5 * 32 * V2 * sizeof(T);
        if (lean_bytes > 48 * 1024)
            cudaFuncSetAttribute(stepK_skew_umu<T, V2, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_bytes);
        stepK_skew_umu<T, V2, K><<<grid, warps * 32, lean_bytes, st>>>(P, skip, ntx2);
        return cudaGetLastError();
    }
    if (variant == 2) {                                          // stepK_lean on a uniform mu_coef plane (P.mu_uniform)
        constexpr int ring_rows = KRing<K>::ROWS > 0 ? KRing<K>::ROWS : 4;
        const size_t lean_bytes = (size_t)warps * ring_rows * 5 * 32 * V2 * sizeof(T);
        if (lean_bytes > 48 * 1024)
            cudaFuncSetAttribute(stepK_lean_umu<T, V2, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_bytes);
        stepK_lean_umu<T, V2, K><<<grid, warps * 32, lean_bytes, st>>>(P, skip, ntx2);
        return cudaGetLastError();
    }
    if (variant == 1) {                                          // fewer issue slots per row (kernels.cuh stepK_lean)
        constexpr int ring_rows = KRing<K>::ROWS > 0 ? KRing<K>::ROWS : 4;
        const size_t lean_bytes = (size_t)warps * ring_rows * 5 * 32 * V2 * sizeof(T);
        if (lean_bytes > 48 * 1024)
            cudaFuncSetAttribute(stepK_lean<T, V2, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)lean_bytes);
        stepK_lean<T, V2, K><<<grid, warps * 32, lean_bytes, st>>>(P, skip, ntx2);
        return cudaGetLastError();
    }
    const size_t ring_bytes = (size_t)warps * KRing<K>::ROWS * 5 * 32 * V2 * sizeof(T);   // per-warp cp.async prefetch ring
    if (ring_bytes > 48 * 1024)
        cudaFuncSetAttribute(stepK_stream<T, V2, K>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)ring_bytes);
    stepK_stream<T, V2, K><<<grid, warps * 32, ring_bytes, st>>>(P, skip, ntx2);
    return cudaGetLastError();
}

#define FDTD_INSTANTIATE(K) \
    template cudaError_t launch_stepK_k<FDTD_PART_T, K>(const StepParams<FDTD_PART_T>&, int, int, dim3, const unsigned char*, int, cudaStream_t);
#if FDTD_K_LO <= 2 && 2 <= FDTD_K_HI
FDTD_INSTANTIATE(2)
#endif
#if FDTD_K_LO <= 3 && 3 <= FDTD_K_HI
FDTD_INSTANTIATE(3)
#endif
#if FDTD_K_LO <= 4 && 4 <= FDTD_K_HI
FDTD_INSTANTIATE(4)
#endif
#if FDTD_K_LO <= 5 && 5 <= FDTD_K_HI
FDTD_INSTANTIATE(5)
#endif
#if FDTD_K_LO <= 6 && 6 <= FDTD_K_HI
FDTD_INSTANTIATE(6)
#endif
#if FDTD_K_LO <= 7 && 7 <= FDTD_K_HI
FDTD_INSTANTIATE(7)
#endif
#if FDTD_K_LO <= 8 && 8 <= FDTD_K_HI
FDTD_INSTANTIATE(8)
#endif
#undef FDTD_INSTANTIATE

}  // namespace fdtd
