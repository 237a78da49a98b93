// Host-callable launchers of the kernels.  Every kernel family is instantiated in its own translation unit (k_*.cu,
// several of them compiled more than once for a slice of the dtype x K space), so that fdtd2d.cu -- context, planner,
// halo transport, run loops, C ABI -- holds no kernel code, the objects build in parallel, and an edit to one family
// rebuilds only that family (fdtd_em_solver_b200/_lib.py caches the objects by content hash).
#pragma once
#include "kernels.cuh"
#include "resident.cuh"
#include "edge.cuh"

namespace fdtd {

// ---- k_step.cu: single-step kernels and the helpers around them ----------------------------------------------------
template <typename T> cudaError_t launch_step_stream(const StepParams<T>& P, int vec, int warps, dim3 grid, cudaStream_t s);
template <typename T> cudaError_t launch_step_exact(const StepParams<T>& P, dim3 grid, cudaStream_t s);
template <typename T> cudaError_t launch_plane_from_f64(dim3 grid, cudaStream_t s, T* dst, long long pitch, const double* src,
                                                        long long src_ld, int rows, int cols);
template <typename T> cudaError_t launch_plane_to_f64(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld,
                                                      const T* src, long long pitch, int rows, int cols, int row_global0,
                                                      const PatchList& patches);
template <typename T> cudaError_t launch_plane_view(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld,
                                                    const T* src, long long pitch, int rows_out, int cols_out, int row_stride,
                                                    int col_stride);
template <typename T> cudaError_t launch_plane_uniform(dim3 grid, cudaStream_t s, const T* src, long long pitch, int rows, int cols,
                                                       T* out, int* differs);
template <typename T> cudaError_t launch_synth_coeffs(dim3 grid, cudaStream_t s, T* epc, T* muc, long long pitch, int lrows,
                                                      int cols, int row_off, int NR, long long NC, unsigned long long seed,
                                                      double dtds, double eps0, double mu0);
template <typename T> cudaError_t launch_paint_coeffs(dim3 grid, cudaStream_t s, T* epc, T* muc, long long pitch, int lrows,
                                                      int cols, int row_off, int NR, const ShapeDev* shapes, int n_shapes,
                                                      double eps_bg, double mu_bg, double dtds);

// ---- k_stepk.cu: the plain K-step kernels (variant 0 = stepK_stream, 1 = stepK_lean), one launcher per (T, K) --------
template <typename T, int K> cudaError_t launch_stepK_k(const StepParams<T>& P, int variant, int warps, dim3 grid,
                                                        const unsigned char* skip, int ntx2, cudaStream_t st);

// ---- k_edge.cu: stepK_edge, `flags` = one of EDGE_CLASSES -----------------------------------------------------------------
template <typename T, int K> cudaError_t launch_edge_k(const StepParams<T>& P, int flags, const EdgeTile* tiles, int n,
                                                       const ResSources<T>& R, long long first, cudaStream_t st);

// ---- k_split.cu: split field storage (FDTD_F32X) -- stepK_edge<double, 2, K, EDGE_ALL, split> for K = 1 .. 8, the plain
// K-step kernel stepK_lean_split<K> for K = 2 .. 8, conversions ----
template <int K> cudaError_t launch_edge_split_k(const StepParams<double>& P, const EdgeTile* tiles, int n,
                                                 const ResSources<double>& R, long long first, cudaStream_t st);
template <int K> cudaError_t launch_stepK_split_k(const StepParams<double>& P, int warps, dim3 grid, const unsigned char* skip,
                                                  int ntx2, cudaStream_t st);
cudaError_t launch_split_from_f64(dim3 grid, cudaStream_t s, float* hi, unsigned short* lo, long long pitch, const double* src,
                                  long long src_ld, int rows, int cols);
cudaError_t launch_split_to_f64(dim3 grid, dim3 block, cudaStream_t s, double* dst, long long dst_ld, const float* hi,
                                const unsigned short* lo, long long pitch, int rows_out, int cols_out, int row_stride,
                                int col_stride, int row_global0, const PatchList& patches);

// ---- k_resident.cu: step_resident (cooperative) ---------------------------------------------------------------------------
template <typename T> cudaError_t resident_blocks_per_sm(int* per_sm);
template <typename T> cudaError_t launch_step_resident(const StepParams<T>& P, T* h1, T* ex1, T* ey1, const ResSources<T>& R,
                                                       long long first, int ns, int blocks, cudaStream_t st);

}  // namespace fdtd
