// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or reachable from the package.
//
// CPU emulation of frame_window (fdtd_em_solver_b200/csrc/window.cuh): the kernel's own __host__ __device__ phase
// functions, driven by plain loops instead of CUDA threads, so that tests/test_window_emulation.py can pin the K-call
// window algorithm (apron K+1, phases A..E, h/hn swap, per-call pulse records, probes, tile-only stores) bitwise to the
// oracle without a GPU.  margin_override: -1 = the kernel's rule (call m runs on the window shrunk by window_margin(m)
// cells), otherwise the shrink per call to try instead (0 = never shrink, 2 = too greedy).  Cell order inside a phase: 0 ascending, 1 descending, 2 a fixed pseudo-random permutation
// (the kernel's threads run a phase's cells in no particular order, so the result must not depend on it).
#include "window.cuh"

#include <cstdint>
#include <numeric>
#include <vector>

using namespace fdtd;

extern "C" {

struct WinEmuSource { int32_t kind, row, col, pad; double mag, magz; };   // = SrcStep<double>

struct WinEmuScene {
    double* gen_a[3];           // h, ex, ey of generation A: lrows x pitch each (local rows, row_off = first stored row)
    double* gen_b[3];
    const double* muc; const double* epc;
    int64_t pitch;
    int32_t NR, NC;
    int32_t row_off, lrows;     // rows this context stores (y-slab) -- whole grid: 0, NR+1
    int32_t row_lo, row_hi;     // rows of the (NR+1)-row index space the frame-tile grid spans
    int32_t absorb[4];
    double S, oms;
    int32_t n_rect;
    int32_t rh[FDTD_MAX_RECT][4], rex[FDTD_MAX_RECT][4], rey[FDTD_MAX_RECT][4];
    int32_t n_probe;
    int32_t probe_ij[FDTD_MAX_PROBE][2];
    double* probe_out;          // K x n_probe x 3
    const int32_t* count; const double* last; const WinEmuSource* src; int32_t stride;
};

int window_emulate_f64(const WinEmuScene* sc, int n_tiles, const int32_t* tiles_xy, int tile_rows, int tile_cols,
                       int K, int apron_override, int order, long long first_step, int margin_override) {
    static_assert(sizeof(WinEmuSource) == sizeof(SrcStep<double>), "record layout");
    if (!sc || K < 1 || K > 8 || order < 0 || order > 2) return -1;
    StepParams<double> P{};
    P.h = sc->gen_a[0]; P.ex = sc->gen_a[1]; P.ey = sc->gen_a[2];
    P.hn = sc->gen_b[0]; P.exn = sc->gen_b[1]; P.eyn = sc->gen_b[2];
    P.muc = sc->muc; P.epc = sc->epc; P.pitch = sc->pitch; P.NR = sc->NR; P.NC = sc->NC;
    P.row_off = sc->row_off; P.lrows = sc->lrows; P.row_lo = sc->row_lo; P.row_hi = sc->row_hi; P.tile_rows = tile_rows;
    P.absorb_u = sc->absorb[0]; P.absorb_d = sc->absorb[1]; P.absorb_l = sc->absorb[2]; P.absorb_r = sc->absorb[3];
    P.S = sc->S; P.oms = sc->oms;
    P.n_rect = sc->n_rect;
    for (int r = 0; r < sc->n_rect; ++r) {
        P.rh[r] = RectI{sc->rh[r][0], sc->rh[r][1], sc->rh[r][2], sc->rh[r][3]};
        P.rex[r] = RectI{sc->rex[r][0], sc->rex[r][1], sc->rex[r][2], sc->rex[r][3]};
        P.rey[r] = RectI{sc->rey[r][0], sc->rey[r][1], sc->rey[r][2], sc->rey[r][3]};
    }
    P.n_probe = sc->n_probe;
    for (int k = 0; k < sc->n_probe; ++k) { P.probe_i[k] = sc->probe_ij[k][0]; P.probe_j[k] = sc->probe_ij[k][1]; }
    P.probe_out = sc->probe_out;
    ResSources<double> R{sc->count, sc->last, reinterpret_cast<const SrcStep<double>*>(sc->src), sc->stride};

    const int apron = apron_override > 0 ? apron_override : window_apron(K);
    Win<double> W;
    W.WR = tile_rows + 2 * apron; W.WC = tile_cols + 2 * apron;
    const int cells = W.WR * W.WC;
    std::vector<double> buf[6];
    for (auto& b : buf) b.assign(cells, 0.0);
    std::vector<int> perm(cells);
    std::iota(perm.begin(), perm.end(), 0);
    if (order == 1) for (int k = 0; k < cells; ++k) perm[k] = cells - 1 - k;
    if (order == 2) {
        uint64_t x = 0x2545F4914F6CDD1Dull;
        for (int k = cells - 1; k > 0; --k) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            std::swap(perm[k], perm[(int)(x % (uint64_t)(k + 1))]);
        }
    }
    std::vector<StepRec<double>> rec(K);
    for (int m = 0; m < K; ++m) window_load_record<double>(rec[m], R, first_step + m);

    for (int t = 0; t < n_tiles; ++t) {
        const int ta = P.row_lo + tiles_xy[2 * t + 1] * tile_rows, ja = tiles_xy[2 * t] * tile_cols;
        const int tb = std::min(ta + tile_rows, P.row_hi), jb = std::min(ja + tile_cols, P.NC + 1);
        W.h = buf[0].data(); W.hn = buf[1].data(); W.ex = buf[2].data(); W.ey = buf[3].data(); W.mu = buf[4].data(); W.ep = buf[5].data();
        W.wa = ta - apron; W.ca = ja - apron;
        for (int k = 0; k < cells; ++k)
            window_put<double>(W, perm[k], window_fetch<double>(P, W.wa + perm[k] / W.WC, W.ca + perm[k] % W.WC));
        const bool ud = window_has_ud_wall<double>(P, W), lr = window_has_lr_wall<double>(P, W);
        for (int m = 0; m < K; ++m) {
            double* const po = P.probe_out ? P.probe_out + (size_t)m * P.n_probe * 3 : nullptr;
            const int mg = margin_override >= 0 ? margin_override * m : window_margin(m);      // cells outside the box are skipped
#define WIN_PHASE(PH, ST, PO) for (int k = 0; k < cells; ++k) { \
        const int li = perm[k] / W.WC, lj = perm[k] % W.WC; \
        if (li >= mg && li < W.WR - mg && lj >= mg && lj < W.WC - mg) \
            window_cell<double, PH>(P, rec[m], W, li, lj, ta, tb, ja, jb, ST, PO); }
            WIN_PHASE('A', false, nullptr) WIN_PHASE('B', false, nullptr)
            if (ud) { WIN_PHASE('C', false, nullptr) }
            if (lr) { WIN_PHASE('D', false, nullptr) }
            WIN_PHASE('E', m == K - 1, po)
#undef WIN_PHASE
            std::swap(W.h, W.hn);
        }
    }
    return 0;
}

}  // extern "C"
