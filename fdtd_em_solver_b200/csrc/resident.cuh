// step_resident: MANY leapfrog steps per launch for grids that live in L2 (the reference's own scenario scripts,
// N = 139 .. 442: tutorial / doubleslit / lens / waveguide).
//
// On those grids one step is a few hundred nanoseconds of work, and the single-step path spends its 43-62 us per step on
// the launch and on the one CTA that walks the closed-form recursion of a corner cell.  Here ONE cooperative launch runs
// n_steps calls of Solver.update (/root/reference/solver.py:169-257): the fields ping-pong between two generations in
// global memory (L2-resident at these sizes), the per-step pulse magnitudes come from a device-resident table, and the
// steps are separated by grid.sync().
//
// Per step every CTA takes (TR x TC)-cell tiles of the (NR+1) x (NC+1) index space.  For a tile it loads the window
// "tile + 2 cells of apron" of the OLD h, ex, ey and the coefficients into shared memory and then executes the
// reference's statement sequence ON THE WINDOW, one phase per statement group, with every predicate (walls, pulses,
// reflectors) evaluated in global coordinates:
//
//   phase 0  load; hard point-source writes into the old h                         solver.py:176-180
//   phase 1  hn = curl, TF/SF fix-ups in list order, reflector zeroing             :182, :184-200, :203-204
//   phase 2  ex, ey in place; TF/SF "right" re-application on ey                   :207-208, :211-217
//   phase 3  up / down absorbing walls on ex, ey (they read phase-2 values)        :220-227
//   phase 4  left / right walls (they read phase-3 values), reflector E zeroing,
//            the h walls (all from the old h, later walls win), store, probes     :228-242, :246-255
//
// Every value of a call depends on old values at most 2 cells away, so every cell of the tile
// sees exactly the operands the reference sees; what the phases compute on the apron from incomplete neighbourhoods is
// never read by a value that is stored.  Within a phase a cell writes only itself and reads only cells no cell writes
// in that phase, so the phases are race-free with one __syncthreads() between them.  Same operations, same order,
// one rounding each => bit-identical to the single-step kernels and to the reference.
//
// The per-cell phase functions are __host__ __device__: tests/host_emul/ compiles them for the CPU and checks the
// window algorithm against the oracle without a GPU (test infrastructure; the package never calls it).
//
// Not handled here (the host keeps such scenes on the single-step path, fdtd2d.cu resident_eligible): y-slabs, and a
// TF/SF "right" line at column 0 (python's h[:, col-1] wraps to the far column there).
#pragma once
#include "kernels.cuh"

#if defined(__CUDACC__)
#include <cooperative_groups.h>
#define FDTD_HD __host__ __device__ __forceinline__
#else
#define FDTD_HD inline
#endif

namespace fdtd {

// arithmetic: the device uses the explicit round-to-nearest intrinsics of kernels.cuh; the host build (emulation test)
// uses the plain IEEE operators and is compiled with -ffp-contract=off
template <typename T> FDTD_HD T r_add(T a, T b) {
#if defined(__CUDA_ARCH__)
    return add(a, b);
#else
    return a + b;
#endif
}
template <typename T> FDTD_HD T r_sub(T a, T b) {
#if defined(__CUDA_ARCH__)
    return sub(a, b);
#else
    return a - b;
#endif
}
template <typename T> FDTD_HD T r_mul(T a, T b) {
#if defined(__CUDA_ARCH__)
    return mul(a, b);
#else
    return a * b;
#endif
}
// fields are rewritten by other CTAs between two grid.sync(): read them through L2 (ld.global.cg), never through
// the non-coherent path; the coefficients are constant for the whole launch
template <typename T> FDTD_HD T r_ld_field(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldcg(p);
#else
    return *p;
#endif
}
template <typename T> FDTD_HD T r_ld_const(const T* p) {
#if defined(__CUDA_ARCH__)
    return __ldg(p);
#else
    return *p;
#endif
}
FDTD_HD bool r_inside(const RectI& r, int i, int j) { return i >= r.r0 && i < r.r1 && j >= r.c0 && j < r.c1; }

template <int TR_, int TC_>
struct ResTile {
    static constexpr int TR = TR_, TC = TC_;       // cells of the index space a tile owns
    static constexpr int APRON = 2;                // dependency radius of one update() call
    static constexpr int WR = TR + 2 * APRON, WC = TC + 2 * APRON;
    static constexpr int CELLS = WR * WC;
};

// the window of one tile: WR x WC cells, local (0,0) = global (wa, ca).  Shared memory on the device.
template <typename T>
struct ResWindow {
    T* h;      // old h incl. the point-source writes (the reference's h_prev)
    T* hn;     // new h before the walls
    T* ex;     // updated in place, like the reference's arrays
    T* ey;
    T* mu;
    T* ep;
    int wa, ca;
};

// per-step pulse records on the device: record n holds the pulses ACTIVE in call n, in list order
template <typename T>
struct ResSources {
    const int* count;          // [table_len]
    const T* last;             // [table_len] magnitude of the last active pulse in list order (App. A-Q5)
    const SrcStep<T>* src;     // [table_len][stride]
    int stride;
};

// what phase 0 brings into the window for one cell: the OLD values, zeros outside the arrays, h already carrying the
// hard point-source writes of solver.py:178-180 (later pulses win)
template <typename T> struct ResCell { T h, mu, ep, ex, ey; };

template <typename T>
FDTD_HD ResCell<T> resident_fetch(const StepParams<T>& P, int i, int j) {
    const bool in_h = i >= 0 && i < P.NR && j >= 0 && j < P.NC;
    const bool in_ex = i >= 0 && i <= P.NR && j >= 0 && j < P.NC;
    const bool in_ey = i >= 0 && i < P.NR && j >= 0 && j <= P.NC;
    const long long o = (long long)i * P.pitch + j;
    ResCell<T> v;
    v.h = v.mu = v.ep = v.ex = v.ey = T(0);
    if (in_h) { v.h = r_ld_field(P.h + o); v.mu = r_ld_const(P.muc + o); v.ep = r_ld_const(P.epc + o); }
    if (in_ex) v.ex = r_ld_field(P.ex + o);
    if (in_ey) v.ey = r_ld_field(P.ey + o);
    if (in_h)
        for (int s = 0; s < P.n_src; ++s)
            if (P.src[s].kind == 0 && P.src[s].row == i && P.src[s].col == j) v.h = P.src[s].mag;
    return v;
}

template <typename T>
FDTD_HD void resident_put(const ResWindow<T>& W, int l, const ResCell<T>& v) {
    W.h[l] = v.h; W.mu[l] = v.mu; W.ep[l] = v.ep; W.ex[l] = v.ex; W.ey[l] = v.ey;
}

template <typename T, typename G, int PHASE>
FDTD_HD void resident_cell(const StepParams<T>& P, const ResWindow<T>& W, int ta, int tb, int ja, int jb, int li, int lj) {
    constexpr int WR = G::WR, WC = G::WC;
    const int i = W.wa + li, j = W.ca + lj;
    const int l = li * WC + lj;
    const int NR = P.NR, NC = P.NC;
    const bool in_h = i >= 0 && i < NR && j >= 0 && j < NC;
    const bool in_ex = i >= 0 && i <= NR && j >= 0 && j < NC;
    const bool in_ey = i >= 0 && i < NR && j >= 0 && j <= NC;
    const bool has_dn = li + 1 < WR, has_up = li >= 1, has_rt = lj + 1 < WC, has_lf = lj >= 1;

    if constexpr (PHASE == 0) {
        resident_put<T>(W, l, resident_fetch<T>(P, i, j));
    } else if constexpr (PHASE == 1) {
        T v = T(0);
        if (in_h && has_dn && has_rt) {
            const T ho = W.h[l], mu = W.mu[l];
            const T dex = r_sub(W.ex[l + WC], W.ex[l]);
            const T eyc = W.ey[l], eyr = W.ey[l + 1];
            v = r_add(ho, r_mul(mu, r_sub(dex, r_sub(eyr, eyc))));              // :182
            for (int s = 0; s < P.n_src; ++s) {                                 // :184-200, list order
                const SrcStep<T>& q = P.src[s];
                if (q.kind == 1) { if (j == q.col) v = r_add(ho, r_mul(mu, r_sub(dex, r_sub(eyr, r_add(eyc, q.magz))))); }
                else if (q.kind == 2) { if (j == q.col) v = r_add(v, r_mul(mu, P.last_mag)); }
                else if (q.kind == 3) { if (i == q.row + 1) v = r_sub(v, r_mul(mu, P.last_mag)); }
                else if (q.kind == 4) { if (i == q.row) v = r_sub(v, r_mul(mu, P.last_mag)); }
            }
            for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rh[r], i, j)) v = T(0);   // :203-204
        }
        W.hn[l] = v;
    } else if constexpr (PHASE == 2) {
        if (in_ex && i >= 1 && i <= NR - 1 && has_up && has_dn && has_rt)       // :207, rows 0 and NR untouched
            W.ex[l] = r_add(W.ex[l], r_mul(W.ep[l], r_sub(W.hn[l], W.hn[l - WC])));
        if (in_ey && has_lf && has_dn && has_rt) {
            T e = W.ey[l];
            if (j >= 1 && j <= NC - 1) e = r_sub(e, r_mul(W.ep[l], r_sub(W.hn[l], W.hn[l - 1])));      // :208
            for (int s = 0; s < P.n_src; ++s)                                   // :211-217 on the ALREADY updated column
                if (P.src[s].kind == 1 && P.src[s].col == j && j >= 1)
                    e = r_sub(e, r_mul(W.ep[l], r_sub(r_sub(W.hn[l], P.src[s].mag), W.hn[l - 1])));
            W.ey[l] = e;
        }
    } else if constexpr (PHASE == 3) {
        if (P.absorb_u && i == 0 && has_dn) {                                   // :222-223
            if (in_ex) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l + WC], P.S));
            if (in_ey) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l + WC], P.S));
        }
        if (P.absorb_d && has_up) {                                             // :226-227 (ex[-1] is row NR, ey[-1] row NR-1)
            if (in_ex && i == NR) W.ex[l] = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l - WC], P.S));
            if (in_ey && i == NR - 1) W.ey[l] = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l - WC], P.S));
        }
    } else {
        if (i < ta || i >= tb || j < ja || j >= jb) return;                     // only the tile's own cells are stored
        T hv = T(0), exv = T(0), eyv = T(0);
        const long long o = (long long)i * P.pitch + j;
        if (in_ex) {
            exv = W.ex[l];
            if (P.absorb_l && j == 0) exv = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l + 1], P.S));            // :230
            if (P.absorb_r && j == NC - 1) exv = r_add(r_mul(W.ex[l], P.oms), r_mul(W.ex[l - 1], P.S));       // :234
            for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rex[r], i, j)) exv = T(0);                      // :238-241
            P.exn[o] = exv;
        }
        if (in_ey) {
            eyv = W.ey[l];
            if (P.absorb_l && j == 0) eyv = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l + 1], P.S));            // :231
            if (P.absorb_r && j == NC) eyv = r_add(r_mul(W.ey[l], P.oms), r_mul(W.ey[l - 1], P.S));           // :235
            for (int r = 0; r < P.n_rect; ++r) if (r_inside(P.rey[r], i, j)) eyv = T(0);                      // :238-242
            P.eyn[o] = eyv;
        }
        if (in_h) {
            hv = W.hn[l];                                                       // walls read the OLD h; later walls win
            if (P.absorb_u && i == 0) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l + WC], P.S));              // :221
            if (P.absorb_d && i == NR - 1) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l - WC], P.S));         // :225
            if (P.absorb_l && j == 0) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l + 1], P.S));               // :229
            if (P.absorb_r && j == NC - 1) hv = r_add(r_mul(W.h[l], P.oms), r_mul(W.h[l - 1], P.S));          // :233
            P.hn[o] = hv;
            for (int k = 0; k < P.n_probe; ++k)                                 // :246-255: raw (h, ex, ey) of the cell
                if (P.probe_i[k] == i && P.probe_j[k] == j) {
                    P.probe_out[3 * k + 0] = (double)hv;
                    P.probe_out[3 * k + 1] = (double)exv;
                    P.probe_out[3 * k + 2] = (double)eyv;
                }
        }
    }
}

// tiles of the (NR+1) x (NC+1) index space
template <typename G> FDTD_HD int resident_tiles_x(int NC) { return (NC + 1 + G::TC - 1) / G::TC; }
template <typename G> FDTD_HD int resident_tiles_y(int NR) { return (NR + 1 + G::TR - 1) / G::TR; }

// the record of call n -> the per-step fields of P (P lives in shared memory on the device)
template <typename T>
FDTD_HD void resident_load_sources(StepParams<T>& P, const ResSources<T>& R, long long n) {
    int cnt = 0;
    T last = T(0);
    if (R.count) { cnt = R.count[n]; last = R.last[n]; }
    for (int s = 0; s < cnt; ++s) P.src[s] = R.src[n * R.stride + s];
    P.n_src = cnt;
    P.last_mag = last;
}

#if defined(__CUDACC__)
constexpr int RES_THREADS = 256;
using ResGeom = ResTile<8, 64>;

template <typename T, typename G, int PHASE>
__device__ __forceinline__ void resident_phase(const StepParams<T>& P, const ResWindow<T>& W, int ta, int tb, int ja, int jb) {
    if constexpr (PHASE == 0) {
        // all of a thread's loads are issued before the first shared-memory store: one L2 round trip per tile, not one
        // per cell (the stores go through generic pointers, so the compiler would not hoist the loads itself)
        constexpr int IT = (G::CELLS + RES_THREADS - 1) / RES_THREADS;
        ResCell<T> v[IT];
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const int idx = threadIdx.x + k * RES_THREADS;
            if (idx < G::CELLS) v[k] = resident_fetch<T>(P, W.wa + idx / G::WC, W.ca + idx % G::WC);
        }
#pragma unroll
        for (int k = 0; k < IT; ++k) {
            const int idx = threadIdx.x + k * RES_THREADS;
            if (idx < G::CELLS) resident_put<T>(W, idx, v[k]);
        }
    } else {
        for (int idx = threadIdx.x; idx < G::CELLS; idx += RES_THREADS)
            resident_cell<T, G, PHASE>(P, W, ta, tb, ja, jb, idx / G::WC, idx % G::WC);
    }
}

// gen0 / gen1: the two generations of (h, ex, ey); call first_step + s reads generation (s & 1 ? gen1 : gen0).
// probe_out of P0 points at the record of the first call of this launch.
template <typename T>
__global__ void __launch_bounds__(RES_THREADS) step_resident(const __grid_constant__ StepParams<T> P0,
                                                             T* h1, T* ex1, T* ey1, const ResSources<T> R,
                                                             long long first_step, int n_steps) {
    using G = ResGeom;
    namespace cg = cooperative_groups;
    cg::grid_group grid = cg::this_grid();
    __shared__ StepParams<T> P;
    __shared__ T sh[G::CELLS], shn[G::CELLS], sex[G::CELLS], sey[G::CELLS], smu[G::CELLS], sep[G::CELLS];
    if (threadIdx.x == 0) P = P0;
    T* const g0[3] = {const_cast<T*>(P0.h), const_cast<T*>(P0.ex), const_cast<T*>(P0.ey)};
    T* const g1[3] = {h1, ex1, ey1};
    const int ntx = resident_tiles_x<G>(P0.NC), nty = resident_tiles_y<G>(P0.NR);
    ResWindow<T> W;
    W.h = sh; W.hn = shn; W.ex = sex; W.ey = sey; W.mu = smu; W.ep = sep;
    for (int s = 0; s < n_steps; ++s) {
        __syncthreads();                                         // everyone is done with the previous step's P
        {   // this call's view of the scene: generations swapped, the call's pulse record (= resident_load_sources,
            // spread over threads so that it costs one memory round trip)
            const long long n = first_step + s;
            const int t = threadIdx.x;
            if (t == 0) {
                const bool odd = s & 1;
                P.h = odd ? g1[0] : g0[0]; P.ex = odd ? g1[1] : g0[1]; P.ey = odd ? g1[2] : g0[2];
                P.hn = odd ? g0[0] : g1[0]; P.exn = odd ? g0[1] : g1[1]; P.eyn = odd ? g0[2] : g1[2];
                P.probe_out = P0.probe_out ? P0.probe_out + (size_t)s * P0.n_probe * 3 : nullptr;
            }
            if (t < R.stride && t < FDTD_MAX_SRC) P.src[t] = R.src[n * R.stride + t];      // entries past count are unused
            if (t == 32) P.n_src = R.count ? R.count[n] : 0;
            if (t == 33) P.last_mag = R.count ? R.last[n] : T(0);
        }
        __syncthreads();
        for (int t = blockIdx.x; t < ntx * nty; t += gridDim.x) {
            const int ta = (t / ntx) * G::TR, ja = (t % ntx) * G::TC;
            const int tb = min(ta + G::TR, P0.NR + 1), jb = min(ja + G::TC, P0.NC + 1);
            W.wa = ta - G::APRON; W.ca = ja - G::APRON;
            resident_phase<T, G, 0>(P, W, ta, tb, ja, jb); __syncthreads();
            resident_phase<T, G, 1>(P, W, ta, tb, ja, jb); __syncthreads();
            resident_phase<T, G, 2>(P, W, ta, tb, ja, jb); __syncthreads();
            resident_phase<T, G, 3>(P, W, ta, tb, ja, jb); __syncthreads();
            resident_phase<T, G, 4>(P, W, ta, tb, ja, jb); __syncthreads();   // the window is reused by the next tile
        }
        if (s + 1 < n_steps) grid.sync();                        // generation B complete and visible before anyone reads it
    }
}
#endif  // __CUDACC__

}  // namespace fdtd
