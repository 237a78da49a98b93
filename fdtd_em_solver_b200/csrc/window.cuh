// frame_window: the FRAME of a temporal-blocking group (walls, pulse lines, reflectors, probes -- every tile the K-step
// kernel stepK_stream must stay away from) advanced K calls in ONE launch, in shared memory.
//
// Today the frame goes through K masked single-step launches (fdtd2d.cu group_impl): K trips through L2 / HBM and K
// dependent launches, ~20 % of a K = 6 group.  Here a CTA takes one frame tile, loads the window "tile + (K+1) cells of
// apron" of generation A once, runs K calls of Solver.update (/root/reference/solver.py:169-257) on the window with the
// phase sequence of step_resident (resident.cuh) -- extended so that a call leaves its results in the window for the
// next call -- and stores the tile's cells of generation B.  It reads only A and writes only frame tiles of B, so it is
// independent of the K-step kernel (which reads A and writes the other tiles of B) and needs no scratch generations.
//
// Phases of one call, one __syncthreads() after each; a cell writes only itself in a phase and reads only cells that
// nobody writes in that phase:
//   A  hard point-source write into the old h (own cell); hn = curl, TF/SF fix-ups, reflector zeroing   solver.py:176-204
//   B  ex, ey in place, TF/SF "right" re-application on ey                                               :207-217
//   C  up / down absorbing walls on ex, ey                                                               :220-227
//   D  left / right absorbing walls on ex, ey                                                            :228-235
//   E  reflector zeroing of ex, ey; h walls (from the old h, later walls win) into hn; probes; on the last call the
//      store of the tile's cells                                                                         :221-242, :246-255
// then h and hn swap roles.  Dependency cone: one cell per call, two inward from a wall on the first hop, so K calls need
// K+1 cells of apron; what the window computes outside the cone of the tile is never read by a stored value.  Call m
// runs on the window shrunk by m cells (window_margin); phases C / D are skipped when the window holds no wall row /
// column.
//
// As in resident.cuh the per-cell functions are __host__ __device__ and tests/host_emul/window_emul.cu pins them bitwise
// to the oracle on the CPU (test infrastructure; never linked into the package).
#pragma once
#include "resident.cuh"

namespace fdtd {

// the pulses ACTIVE in one call, in list order (what fill_params puts into StepParams for a single-step launch)
template <typename T>
struct StepRec {
    int n_src, pad;
    T last_mag;
    SrcStep<T> src[FDTD_MAX_SRC];
};

template <typename T>
struct Win {
    T* h;      // old h of the current call (incl. its point-source writes)
    T* hn;     // new h
    T* ex;
    T* ey;
    T* mu;
    T* ep;
    int wa, ca;       // global index of local (0,0)
    int WR, WC;
};

template <typename T>
FDTD_HD void window_load_record(StepRec<T>& Q, const ResSources<T>& R, long long n) {
    int cnt = 0;
    T last = T(0);
    if (R.count) { cnt = R.count[n]; last = R.last[n]; }
    for (int s = 0; s < cnt; ++s) Q.src[s] = R.src[n * R.stride + s];
    Q.n_src = cnt; Q.pad = 0;
    Q.last_mag = last;
}

// old values of one window cell; zeros outside the arrays and outside the rows this context stores.  COHERENT: the
// fields are rewritten by other CTAs during the launch (resident_window), so they are read through L2.
template <typename T, bool COHERENT = false>
FDTD_HD ResCell<T> window_fetch(const StepParams<T>& P, int i, int j) {
    const bool stored = i >= P.row_off && i < P.row_off + P.lrows;
    const bool in_h = stored && i >= 0 && i < P.NR && j >= 0 && j < P.NC;
    const bool in_ex = stored && i >= 0 && i <= P.NR && j >= 0 && j < P.NC;
    const bool in_ey = stored && i >= 0 && i < P.NR && j >= 0 && j <= P.NC;
    const long long o = (long long)(i - P.row_off) * P.pitch + j;
    ResCell<T> v;
    v.h = v.mu = v.ep = v.ex = v.ey = T(0);
    if (in_h) { v.h = COHERENT ? r_ld_field(P.h + o) : r_ld_const(P.h + o); v.mu = r_ld_const(P.muc + o); v.ep = r_ld_const(P.epc + o); }
    if (in_ex) v.ex = COHERENT ? r_ld_field(P.ex + o) : r_ld_const(P.ex + o);
    if (in_ey) v.ey = COHERENT ? r_ld_field(P.ey + o) : r_ld_const(P.ey + o);
    return v;
}

template <typename T>
FDTD_HD void window_put(const Win<T>& W, int l, const ResCell<T>& v) {
    W.h[l] = v.h; W.hn[l] = T(0); W.mu[l] = v.mu; W.ep[l] = v.ep; W.ex[l] = v.ex; W.ey[l] = v.ey;
}

// One phase of one call for one window cell.  Phase 'E' also takes the tile (cells that are stored / probed) and whether
// this is the last call of the launch; probe_out points at this call's record.
template <typename T, char PHASE>
FDTD_HD void window_cell(const StepParams<T>& P, const StepRec<T>& Q, const Win<T>& W, int li, int lj,
                         int ta = 0, int tb = 0, int ja = 0, int jb = 0, bool store = false, double* probe_out = nullptr) {
    const int WC = W.WC;
    const int i = W.wa + li, j = W.ca + lj;
    const int l = li * WC + lj;
    const int NR = P.NR, NC = P.NC;
    const bool in_h = i >= 0 && i < NR && j >= 0 && j < NC;
    const bool in_ex = i >= 0 && i <= NR && j >= 0 && j < NC;
    const bool in_ey = i >= 0 && i < NR && j >= 0 && j <= NC;
    const bool has_dn = li + 1 < W.WR, has_up = li >= 1, has_rt = lj + 1 < WC, has_lf = lj >= 1;

    if constexpr (PHASE == 'A') {
        T v = T(0);
        if (in_h) {
            T ho = W.h[l];
            for (int s = 0; s < Q.n_src; ++s)                                  // :178-180, later pulses win
                if (Q.src[s].kind == 0 && Q.src[s].row == i && Q.src[s].col == j) ho = Q.src[s].mag;
            W.h[l] = ho;
            if (has_dn && has_rt) {
                const T mu = W.mu[l];
                const T dex = r_sub(W.ex[l + WC], W.ex[l]);
                const T eyc = W.ey[l], eyr = W.ey[l + 1];
                v = r_add(ho, r_mul(mu, r_sub(dex, r_sub(eyr, eyc))));          // :182
                for (int s = 0; s < Q.n_src; ++s) {                             // :184-200, list order
                    const SrcStep<T>& q = Q.src[s];
                    if (q.kind == 1) { if (j == q.col) v = r_add(ho, r_mul(mu, r_sub(dex, r_sub(eyr, r_add(eyc, q.magz))))); }
                    else if (q.kind == 2) { if (j == q.col) v = r_add(v, r_mul(mu, Q.last_mag)); }
                    else if (q.kind == 3) { if (i == q.row + 1) v = r_sub(v, r_mul(mu, Q.last_mag)); }
                    else if (q.kind == 4) { if (i == q.row) v = r_sub(v, r_mul(mu, Q.last_mag)); }
                }
                for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rh[r], i, j)) v = T(0);   // :203-204
            }
        }
        W.hn[l] = v;
    } else if constexpr (PHASE == 'B') {
        if (in_ex && i >= 1 && i <= NR - 1 && has_up && has_dn && has_rt)       // :207, rows 0 and NR untouched
            W.ex[l] = r_add(W.ex[l], r_mul(W.ep[l], r_sub(W.hn[l], W.hn[l - WC])));
        if (in_ey && has_lf && has_dn && has_rt) {
            T e = W.ey[l];
            if (j >= 1 && j <= NC - 1) e = r_sub(e, r_mul(W.ep[l], r_sub(W.hn[l], W.hn[l - 1])));      // :208
            for (int s = 0; s < Q.n_src; ++s)                                   // :211-217 on the ALREADY updated column
                if (Q.src[s].kind == 1 && Q.src[s].col == j && j >= 1)
                    e = r_sub(e, r_mul(W.ep[l], r_sub(r_sub(W.hn[l], Q.src[s].mag), W.hn[l - 1])));
            W.ey[l] = e;
        }
    } else if constexpr (PHASE == 'C') {
        if (P.absorb_u && i == 0 && has_dn) {                                   // :222-223
            if (in_ex) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l + WC], P.S));
            if (in_ey) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l + WC], P.S));
        }
        if (P.absorb_d && has_up) {                                             // :226-227
            if (in_ex && i == NR) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l - WC], P.S));
            if (in_ey && i == NR - 1) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l - WC], P.S));
        }
    } else if constexpr (PHASE == 'D') {
        if (P.absorb_l && j == 0 && has_rt) {                                   // :230-231
            if (in_ex) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l + 1], P.S));
            if (in_ey) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l + 1], P.S));
        }
        if (P.absorb_r && has_lf) {                                             // :234-235
            if (in_ex && j == NC - 1) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l - 1], P.S));
            if (in_ey && j == NC) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l - 1], P.S));
        }
    } else {
        T hv = T(0), exv = T(0), eyv = T(0);
        if (in_ex) {
            exv = W.ex[l];
            for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rex[r], i, j)) exv = T(0);                      // :238-241
            W.ex[l] = exv;
        }
        if (in_ey) {
            eyv = W.ey[l];
            for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rey[r], i, j)) eyv = T(0);                      // :238-242
            W.ey[l] = eyv;
        }
        if (in_h) {
            hv = W.hn[l];                                                       // walls read the OLD h; later walls win
            if (P.absorb_u && i == 0 && has_dn) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l + WC], P.S));            // :221
            if (P.absorb_d && i == NR - 1 && has_up) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l - WC], P.S));       // :225
            if (P.absorb_l && j == 0 && has_rt) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l + 1], P.S));             // :229
            if (P.absorb_r && j == NC - 1 && has_lf) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l - 1], P.S));        // :233
            W.hn[l] = hv;
        }
        if (i < ta || i >= tb || j < ja || j >= jb) return;                     // the rest concerns the tile's own cells
        if (in_h && probe_out)
            for (int k = 0; k < P.n_probe; ++k)                                 // :246-255: raw (h, ex, ey) of the cell
                if (P.probe_i[k] == i && P.probe_j[k] == j) {
                    probe_out[3 * k + 0] = (double)hv;
                    probe_out[3 * k + 1] = (double)exv;
                    probe_out[3 * k + 2] = (double)eyv;
                }
        if (store) {
            const long long o = (long long)(i - P.row_off) * P.pitch + j;
            if (in_h) P.hn[o] = hv;
            if (in_ex) P.exn[o] = exv;
            if (in_ey) P.eyn[o] = eyv;
        }
    }
}

FDTD_HD int window_apron(int K) { return K + 1; }

// Call m (0-based) of a launch only has to be exact on the tile + (K-m) cells: a cell at distance d from the tile reads
// cells at distance <= max(d+1, 2) (one cell per call; two inward from a wall, which exceeds d+1 only for d = 0).  So
// call m runs on the window shrunk by m cells on every side; what lies outside goes stale and is never read by a value
// that matters.
FDTD_HD int window_margin(int m) { return m; }

// Phases C / D do something only if the window holds a row / column an absorbing wall writes.
template <typename T>
FDTD_HD bool window_has_ud_wall(const StepParams<T>& P, const Win<T>& W) {
    const int r0 = W.wa, r1 = W.wa + W.WR;                     // window rows [r0, r1)
    return (P.absorb_u && r0 <= 0 && r1 > 0) || (P.absorb_d && r0 <= P.NR && r1 > P.NR - 1);
}
template <typename T>
FDTD_HD bool window_has_lr_wall(const StepParams<T>& P, const Win<T>& W) {
    const int c0 = W.ca, c1 = W.ca + W.WC;
    return (P.absorb_l && c0 <= 0 && c1 > 0) || (P.absorb_r && c0 <= P.NC && c1 > P.NC - 1);
}

#if defined(__CUDACC__)
constexpr int WIN_THREADS = 256;
constexpr int WIN_MAX_CALLS = 8;

// one phase of call m over the window shrunk by `margin` cells: warps take rows, lanes take columns (no division)
template <typename T, char PHASE>
__device__ __forceinline__ void window_phase(const StepParams<T>& P, const StepRec<T>& Q, const Win<T>& W, int margin,
                                             int ta, int tb, int ja, int jb, bool store, double* probe_out) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int li = margin + warp; li < W.WR - margin; li += WIN_THREADS / 32)
        for (int lj = margin + lane; lj < W.WC - margin; lj += 32)
            window_cell<T, PHASE>(P, Q, W, li, lj, ta, tb, ja, jb, store, probe_out);
    __syncthreads();
}

// K pulse records into shared memory in one memory round trip: thread m*16+s copies pulse s of call m, thread 128+m its
// header.  The caller synchronises before the records are read.
template <typename T>
__device__ __forceinline__ void window_load_records(StepRec<T>* rec, const ResSources<T>& R, long long first_step, int K) {
    const int t = threadIdx.x;
    const int m = t / FDTD_MAX_SRC, s = t % FDTD_MAX_SRC;
    if (m < K && s < R.stride) rec[m].src[s] = R.src[(first_step + m) * R.stride + s];
    if (t >= 128 && t - 128 < K) {
        rec[t - 128].n_src = R.count ? R.count[first_step + t - 128] : 0;
        rec[t - 128].last_mag = R.count ? R.last[first_step + t - 128] : T(0);
    }
}

// One tile: load the window (tile + apron) of P's generation A, run K calls on it, store the tile into generation B.
// probe0 = probe record of the first of the K calls (or null).
template <typename T, bool COHERENT>
__device__ __forceinline__ void window_tile(const StepParams<T>& P, const StepRec<T>* rec, T* base, int ta, int tb, int ja,
                                            int jb, int tile_rows, int tile_cols, int apron, int K, double* probe0) {
    Win<T> W;
    W.WR = tile_rows + 2 * apron; W.WC = tile_cols + 2 * apron;
    W.wa = ta - apron; W.ca = ja - apron;
    const int cells = W.WR * W.WC;
    W.h = base; W.hn = base + cells; W.ex = base + 2 * cells; W.ey = base + 3 * cells; W.mu = base + 4 * cells; W.ep = base + 5 * cells;
    for (int b = 0; b < cells; b += 4 * WIN_THREADS) {          // four cells per thread and memory round trip
        ResCell<T> v[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = b + k * WIN_THREADS + threadIdx.x;
            if (idx < cells) v[k] = window_fetch<T, COHERENT>(P, W.wa + idx / W.WC, W.ca + idx % W.WC);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const int idx = b + k * WIN_THREADS + threadIdx.x;
            if (idx < cells) window_put<T>(W, idx, v[k]);
        }
    }
    __syncthreads();
    const bool ud = window_has_ud_wall<T>(P, W), lr = window_has_lr_wall<T>(P, W);
    for (int m = 0; m < K; ++m) {
        double* const probe_out = probe0 ? probe0 + (size_t)m * P.n_probe * 3 : nullptr;
        const int mg = window_margin(m);
        window_phase<T, 'A'>(P, rec[m], W, mg, ta, tb, ja, jb, false, nullptr);
        window_phase<T, 'B'>(P, rec[m], W, mg, ta, tb, ja, jb, false, nullptr);
        if (ud) window_phase<T, 'C'>(P, rec[m], W, mg, ta, tb, ja, jb, false, nullptr);
        if (lr) window_phase<T, 'D'>(P, rec[m], W, mg, ta, tb, ja, jb, false, nullptr);
        window_phase<T, 'E'>(P, rec[m], W, mg, ta, tb, ja, jb, m == K - 1, probe_out);
        T* const t = W.h; W.h = W.hn; W.hn = t;                  // the new h is the next call's old h
    }
}

// tiles[b] = (column tile, row tile) of the frame-tile grid: rows P0.row_lo + ty*tile_rows .., cols tx*tile_cols ..
// P0: generation A in h/ex/ey, generation B in hn/exn/eyn, probe_out = record of call first_step.
template <typename T>
__global__ void __launch_bounds__(WIN_THREADS) frame_window(const __grid_constant__ StepParams<T> P0,
                                                            const int2* __restrict__ tiles, int tile_rows, int tile_cols,
                                                            int K, const ResSources<T> R, long long first_step) {
    extern __shared__ __align__(16) unsigned char win_raw[];
    __shared__ StepRec<T> rec[WIN_MAX_CALLS];
    const int2 tile = tiles[blockIdx.x];
    const int ta = P0.row_lo + tile.y * tile_rows, ja = tile.x * tile_cols;
    const int tb = min(ta + tile_rows, P0.row_hi), jb = min(ja + tile_cols, P0.NC + 1);
    window_load_records<T>(rec, R, first_step, K);
    window_tile<T, false>(P0, rec, reinterpret_cast<T*>(win_raw), ta, tb, ja, jb, tile_rows, tile_cols, window_apron(K), K,
                          P0.probe_out);
}

// Resident mode with M calls per grid barrier (the many-calls-per-launch kernel of resident.cuh, built from the K-call
// window): every CTA strides over the tiles of the whole index space; per interval of M calls it loads tile + M+1 cells
// of apron, runs the M calls in shared memory and stores the tile into the other generation, then the grid synchronises.
// One barrier, one window load and one store per M calls instead of per call, for (tile + 2(M+1))^2 / tile work.
// P0: generation 0 in h/ex/ey (read by the first interval), generation 1 in hn/exn/eyn.
template <typename T>
__global__ void __launch_bounds__(WIN_THREADS) resident_window(const __grid_constant__ StepParams<T> P0, int tile_rows,
                                                               int tile_cols, int M, const ResSources<T> R,
                                                               long long first_step, int n_steps) {
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    extern __shared__ __align__(16) unsigned char win_raw[];
    __shared__ StepRec<T> rec[WIN_MAX_CALLS];
    __shared__ StepParams<T> P;
    if (threadIdx.x == 0) P = P0;
    const int ntx = (P0.NC + 1 + tile_cols - 1) / tile_cols, nty = (P0.NR + 1 + tile_rows - 1) / tile_rows;
    const int apron = window_apron(M);
    int interval = 0;
    for (int s = 0; s < n_steps; s += M, ++interval) {
        const int K = min(M, n_steps - s);
        __syncthreads();                                         // everyone is done with the previous interval's P and records
        if (threadIdx.x == 0) {
            const bool odd = interval & 1;
            P.h = odd ? P0.hn : P0.h; P.ex = odd ? P0.exn : P0.ex; P.ey = odd ? P0.eyn : P0.ey;
            P.hn = odd ? const_cast<T*>(P0.h) : P0.hn; P.exn = odd ? const_cast<T*>(P0.ex) : P0.exn;
            P.eyn = odd ? const_cast<T*>(P0.ey) : P0.eyn;
        }
        window_load_records<T>(rec, R, first_step + s, K);
        __syncthreads();
        double* const probe0 = P0.probe_out ? P0.probe_out + (size_t)s * P0.n_probe * 3 : nullptr;
        for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
            const int ta = (t / ntx) * tile_rows, ja = (t % ntx) * tile_cols;
            const int tb = min(ta + tile_rows, P0.NR + 1), jb = min(ja + tile_cols, P0.NC + 1);
            window_tile<T, true>(P, rec, reinterpret_cast<T*>(win_raw), ta, tb, ja, jb, tile_rows, tile_cols, apron, K, probe0);
        }
        if (s + M < n_steps) grid.sync();                        // the other generation is complete and visible
    }
}
#endif  // __CUDACC__

}  // namespace fdtd
