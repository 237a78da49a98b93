// stepK_edge (edge.cuh) in its five tile classes for ONE dtype and a range of K: compiled several times with
// -DFDTD_PART_T=double|float -DFDTD_K_LO=.. -DFDTD_K_HI=..  See launch.h.
#include "launch.h"

#ifndef FDTD_PART_T
#error "compile with -DFDTD_PART_T=double|float -DFDTD_K_LO=a -DFDTD_K_HI=b"
#endif

namespace fdtd {

template <typename T, int K>
cudaError_t launch_edge_k(const StepParams<T>& P, int flags, const EdgeTile* tiles, int n, const ResSources<T>& R,
                          long long first, cudaStream_t st) {
    constexpr int V2 = sizeof(T) == 8 ? 2 : 4;
    switch (flags) {
        case EDGE_COLS: stepK_edge<T, V2, K, EDGE_COLS><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first); break;
        case EDGE_ROWS: stepK_edge<T, V2, K, EDGE_ROWS><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first); break;
        case EDGE_SRC: stepK_edge<T, V2, K, EDGE_SRC><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first); break;
        case EDGE_COLS | EDGE_SRC: stepK_edge<T, V2, K, EDGE_COLS | EDGE_SRC><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first); break;
        case EDGE_ALL: stepK_edge<T, V2, K, EDGE_ALL><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first); break;
        default: return cudaErrorInvalidValue;
    }
    return cudaGetLastError();
}

#define FDTD_INSTANTIATE(K) \
    template cudaError_t launch_edge_k<FDTD_PART_T, K>(const StepParams<FDTD_PART_T>&, int, const EdgeTile*, int, \
                                                       const ResSources<FDTD_PART_T>&, long long, cudaStream_t);
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
