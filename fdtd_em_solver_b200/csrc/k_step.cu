// Single-step kernels (step_stream, step_exact) and the helper kernels (fp32 staging, snapshot pack, coefficient
// synthesis / painting) for both dtypes.  See launch.h.
#include "launch.h"

namespace fdtd {

template <typename T>
cudaError_t launch_step_stream(const StepParams<T>& P, int vec, int warps, dim3 grid, cudaStream_t s) {
    constexpr int VA = sizeof(T) == 8 ? 2 : 4, VB = 2 * VA;
    if (vec == VA) step_stream<T, VA><<<grid, warps * 32, 0, s>>>(P);
    else if (vec == VB) step_stream<T, VB><<<grid, warps * 32, 0, s>>>(P);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_step_exact(const StepParams<T>& P, dim3 grid, cudaStream_t s) {
    step_exact<T><<<grid, 256, 0, s>>>(P);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_plane_from_f64(dim3 grid, cudaStream_t s, T* dst, long long pitch, const double* src, long long src_ld,
                                  int rows, int cols) {
    plane_from_f64<T><<<grid, 256, 0, s>>>(dst, pitch, src, src_ld, rows, cols);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_plane_to_f64(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld, const T* src,
                                long long pitch, int rows, int cols, int row_global0, const PatchList& patches) {
    plane_to_f64<T><<<grid, block, 0, s>>>(dst, dst_ld, src, pitch, rows, cols, row_global0, patches);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_plane_view(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld, const T* src,
                              long long pitch, int rows_out, int cols_out, int row_stride, int col_stride) {
    plane_view<T><<<grid, block, 0, s>>>(dst, dst_ld, src, pitch, rows_out, cols_out, row_stride, col_stride);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_plane_uniform(dim3 grid, cudaStream_t s, const T* src, long long pitch, int rows, int cols, T* out, int* differs) {
    plane_uniform<T><<<grid, 256, 0, s>>>(src, pitch, rows, cols, out, differs);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_synth_coeffs(dim3 grid, cudaStream_t s, T* epc, T* muc, long long pitch, int lrows, int cols, int row_off,
                                int NR, long long NC, unsigned long long seed, double dtds, double eps0, double mu0) {
    synth_coeffs<T><<<grid, 256, 0, s>>>(epc, muc, pitch, lrows, cols, row_off, NR, NC, seed, dtds, eps0, mu0);
    return cudaGetLastError();
}

template <typename T>
cudaError_t launch_paint_coeffs(dim3 grid, cudaStream_t s, T* epc, T* muc, long long pitch, int lrows, int cols, int row_off,
                                int NR, const ShapeDev* shapes, int n_shapes, double eps_bg, double mu_bg, double dtds) {
    paint_coeffs<T><<<grid, 256, 0, s>>>(epc, muc, pitch, lrows, cols, row_off, NR, shapes, n_shapes, eps_bg, mu_bg, dtds);
    return cudaGetLastError();
}

#define FDTD_INSTANTIATE(T) \
    template cudaError_t launch_step_stream<T>(const StepParams<T>&, int, int, dim3, cudaStream_t); \
    template cudaError_t launch_step_exact<T>(const StepParams<T>&, dim3, cudaStream_t); \
    template cudaError_t launch_plane_from_f64<T>(dim3, cudaStream_t, T*, long long, const double*, long long, int, int); \
    template cudaError_t launch_plane_to_f64<T>(dim3, dim3, cudaStream_t, double*, long long, const T*, long long, int, int, int, \
                                                const PatchList&); \
    template cudaError_t launch_plane_view<T>(dim3, dim3, cudaStream_t, double*, long long, const T*, long long, int, int, int, int); \
    template cudaError_t launch_plane_uniform<T>(dim3, cudaStream_t, const T*, long long, int, int, T*, int*); \
    template cudaError_t launch_synth_coeffs<T>(dim3, cudaStream_t, T*, T*, long long, int, int, int, int, long long, \
                                                unsigned long long, double, double, double); \
    template cudaError_t launch_paint_coeffs<T>(dim3, cudaStream_t, T*, T*, long long, int, int, int, int, const ShapeDev*, int, \
                                                double, double, double);
FDTD_INSTANTIATE(double)
FDTD_INSTANTIATE(float)
#undef FDTD_INSTANTIATE

}  // namespace fdtd
