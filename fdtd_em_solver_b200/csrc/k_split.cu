// Split field storage (dtype FDTD_F32X: fp32 hi + bf16 lo planes, double arithmetic; kernels.cuh): stepK_edge with every
// feature for K = 1 .. 8 -- in this mode EVERY tile of the grid is an edge tile -- and the conversions between the split
// planes and the float64 arrays of the host.  Compiled per slice of K (-DFDTD_K_LO / -DFDTD_K_HI; the conversions ride
// in the slice that holds K = 1).  See launch.h.
#include "launch.h"

#ifndef FDTD_K_LO
#error "compile with -DFDTD_K_LO=a -DFDTD_K_HI=b"
#endif

namespace fdtd {

template <int K>
cudaError_t launch_edge_split_k(const StepParams<double>& P, const EdgeTile* tiles, int n, const ResSources<double>& R,
                                long long first, cudaStream_t st) {
    stepK_edge<double, 2, K, EDGE_ALL, 1><<<(unsigned)n, 32, 0, st>>>(P, tiles, n, R, first);
    return cudaGetLastError();
}

// the plain tiles of a K-step group under split storage (kernels.cuh stepK_lean_split), K = 2 .. 8
template <int K>
cudaError_t launch_stepK_split_k(const StepParams<double>& P, int warps, dim3 grid, const unsigned char* skip, int ntx2,
                                 cudaStream_t st) {
    const size_t ring_bytes = (size_t)warps * 4 * SplitRing::SLOT;
    stepK_lean_split<K><<<grid, warps * 32, ring_bytes, st>>>(P, skip, ntx2);
    return cudaGetLastError();
}

#define FDTD_INSTANTIATE_EDGE(K) \
    template cudaError_t launch_edge_split_k<K>(const StepParams<double>&, const EdgeTile*, int, const ResSources<double>&, \
                                                long long, cudaStream_t);
#define FDTD_INSTANTIATE(K) FDTD_INSTANTIATE_EDGE(K) \
    template cudaError_t launch_stepK_split_k<K>(const StepParams<double>&, int, dim3, const unsigned char*, int, cudaStream_t);
#if FDTD_K_LO <= 1 && 1 <= FDTD_K_HI
FDTD_INSTANTIATE_EDGE(1)                                         // single calls have no plain tiles (all-edge launch)

// dense float64 rows -> split planes of a pitched buffer (hi at `hi`, lo at `lo`)
__global__ void split_from_f64(float* hi, unsigned short* lo, long long pitch, const double* src, long long src_ld, int rows,
                               int cols) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols) return;
    for (int i = blockIdx.y; i < rows; i += gridDim.y)
        split_encode(src[(long long)i * src_ld + j], hi[(long long)i * pitch + j], lo[(long long)i * pitch + j]);
}

// split planes -> dense float64, every row_stride-th row / col_stride-th column, with the snapshot aliasing patches
// (plane_to_f64's; patch rows and columns are in units of the source array)
__global__ void split_to_f64(double* dst, long long dst_ld, const float* hi, const unsigned short* lo, long long pitch,
                             int rows_out, int cols_out, int row_stride, int col_stride, int row_global0,
                             const __grid_constant__ PatchList patches) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x;
    if (j >= cols_out) return;
    for (int i = blockIdx.y * blockDim.y + threadIdx.y; i < rows_out; i += gridDim.y * blockDim.y) {
        const long long at = (long long)i * row_stride * pitch + (long long)j * col_stride;
        double v = split_decode(hi[at], lo[at]);
        for (int k = 0; k < patches.n; ++k)
            if (patches.p[k].row == row_global0 + i * row_stride && patches.p[k].col == j * col_stride) v = patches.p[k].value;
        dst[(long long)i * dst_ld + j] = v;
    }
}

cudaError_t launch_split_from_f64(dim3 grid, cudaStream_t s, float* hi, unsigned short* lo, long long pitch, const double* src,
                                  long long src_ld, int rows, int cols) {
    split_from_f64<<<grid, 256, 0, s>>>(hi, lo, pitch, src, src_ld, rows, cols);
    return cudaGetLastError();
}

cudaError_t launch_split_to_f64(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld, const float* hi,
                                const unsigned short* lo, long long pitch, int rows_out, int cols_out, int row_stride,
                                int col_stride, int row_global0, const PatchList& patches) {
    split_to_f64<<<grid, block, 0, s>>>(dst, dst_ld, hi, lo, pitch, rows_out, cols_out, row_stride, col_stride, row_global0,
                                        patches);
    return cudaGetLastError();
}
#endif
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
#undef FDTD_INSTANTIATE_EDGE

}  // namespace fdtd
