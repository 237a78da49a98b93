// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or reachable from the package.
//
// CPU emulation of the window algorithm of step_resident (fdtd_em_solver_b200/csrc/resident.cuh): the very same
// __host__ __device__ per-cell phase functions the kernel runs, driven by plain loops instead of CUDA threads, so that
// tests/test_resident_emulation.py can check the tile/apron/phase logic against the oracle in a container without a
// GPU.  What it cannot cover (grid.sync, the cooperative launch, shared-memory sizing) is covered by the -m gpu tests.
//
// Cell order inside a phase: 0 = ascending, 1 = descending, 2 = a fixed pseudo-random permutation.  The kernel's threads
// run a phase's cells in no particular order, so the result must not depend on it -- a cell that read what another
// cell writes in the same phase would show up as a difference between the orders.
#include "resident.cuh"

#include <cstdint>
#include <numeric>
#include <vector>

using namespace fdtd;

extern "C" {

struct EmuSource { int32_t kind, row, col, pad; double mag, magz; };   // = SrcStep<double>

struct EmuScene {
    double* gen0[3];            // h, ex, ey (rows+1) x pitch each, generation read by the first call
    double* gen1[3];
    const double* muc; const double* epc;
    int64_t pitch;
    int32_t NR, NC;
    int32_t absorb[4];          // up, down, left, right
    double S, oms;
    int32_t n_rect;
    int32_t rh[FDTD_MAX_RECT][4], rex[FDTD_MAX_RECT][4], rey[FDTD_MAX_RECT][4];   // already clipped per array
    int32_t n_probe;
    int32_t probe_ij[FDTD_MAX_PROBE][2];
    double* probe_out;          // n_steps x n_probe x 3
    const int32_t* count; const double* last; const EmuSource* src; int32_t stride;   // per-call records (may be NULL)
};

}  // extern "C"

namespace {

template <typename G>
int emulate(const EmuScene& sc, int order, long long first_step, int n_steps) {
    static_assert(sizeof(EmuSource) == sizeof(SrcStep<double>), "record layout");
    StepParams<double> P{};
    P.muc = sc.muc; P.epc = sc.epc; P.pitch = sc.pitch; P.NR = sc.NR; P.NC = sc.NC;
    P.row_off = 0; P.lrows = sc.NR + 1; P.row_lo = 0; P.row_hi = sc.NR + 1; P.tile_rows = G::TR;
    P.absorb_u = sc.absorb[0]; P.absorb_d = sc.absorb[1]; P.absorb_l = sc.absorb[2]; P.absorb_r = sc.absorb[3];
    P.S = sc.S; P.oms = sc.oms;
    P.n_rect = sc.n_rect;
    for (int r = 0; r < sc.n_rect; ++r) {
        P.rh[r] = RectI{sc.rh[r][0], sc.rh[r][1], sc.rh[r][2], sc.rh[r][3]};
        P.rex[r] = RectI{sc.rex[r][0], sc.rex[r][1], sc.rex[r][2], sc.rex[r][3]};
        P.rey[r] = RectI{sc.rey[r][0], sc.rey[r][1], sc.rey[r][2], sc.rey[r][3]};
    }
    P.n_probe = sc.n_probe;
    for (int k = 0; k < sc.n_probe; ++k) { P.probe_i[k] = sc.probe_ij[k][0]; P.probe_j[k] = sc.probe_ij[k][1]; }
    ResSources<double> R{sc.count, sc.last, reinterpret_cast<const SrcStep<double>*>(sc.src), sc.stride};

    std::vector<double> buf[6];
    for (auto& b : buf) b.assign(G::CELLS, 0.0);
    ResWindow<double> W;
    W.h = buf[0].data(); W.hn = buf[1].data(); W.ex = buf[2].data(); W.ey = buf[3].data(); W.mu = buf[4].data(); W.ep = buf[5].data();

    std::vector<int> perm(G::CELLS);
    std::iota(perm.begin(), perm.end(), 0);
    if (order == 1) for (int k = 0; k < G::CELLS; ++k) perm[k] = G::CELLS - 1 - k;
    if (order == 2) {
        uint64_t x = 0x2545F4914F6CDD1Dull;
        for (int k = G::CELLS - 1; k > 0; --k) {
            x ^= x << 13; x ^= x >> 7; x ^= x << 17;
            std::swap(perm[k], perm[(int)(x % (uint64_t)(k + 1))]);
        }
    }
    const int ntx = resident_tiles_x<G>(sc.NC), nty = resident_tiles_y<G>(sc.NR);
    for (int s = 0; s < n_steps; ++s) {
        const bool odd = s & 1;
        double* const* a = odd ? sc.gen1 : sc.gen0;
        double* const* b = odd ? sc.gen0 : sc.gen1;
        P.h = a[0]; P.ex = a[1]; P.ey = a[2]; P.hn = b[0]; P.exn = b[1]; P.eyn = b[2];
        P.probe_out = sc.probe_out ? sc.probe_out + (size_t)s * sc.n_probe * 3 : nullptr;
        resident_load_sources<double>(P, R, first_step + s);
        for (int t = 0; t < ntx * nty; ++t) {
            const int ta = (t / ntx) * G::TR, ja = (t % ntx) * G::TC;
            const int tb = std::min(ta + G::TR, sc.NR + 1), jb = std::min(ja + G::TC, sc.NC + 1);
            W.wa = ta - G::APRON; W.ca = ja - G::APRON;
#define EMU_PHASE(PH) for (int k = 0; k < G::CELLS; ++k) resident_cell<double, G, PH>(P, W, ta, tb, ja, jb, perm[k] / G::WC, perm[k] % G::WC);
            EMU_PHASE(0) EMU_PHASE(1) EMU_PHASE(2) EMU_PHASE(3) EMU_PHASE(4)
#undef EMU_PHASE
        }
    }
    return n_steps & 1;          // generation that holds the result
}

}  // namespace

extern "C" int resident_emulate_f64(const EmuScene* sc, int geom, int order, long long first_step, int n_steps) {
    if (!sc || n_steps < 0 || order < 0 || order > 2) return -1;
    switch (geom) {
        case 0: return emulate<ResTile<8, 64>>(*sc, order, first_step, n_steps);     // the kernel's geometry (ResGeom)
        case 1: return emulate<ResTile<1, 1>>(*sc, order, first_step, n_steps);      // every cell its own tile: the apron is all there is
        case 2: return emulate<ResTile<3, 5>>(*sc, order, first_step, n_steps);
        case 3: return emulate<ResTile<2, 7>>(*sc, order, first_step, n_steps);
        default: return -1;
    }
}

extern "C" int resident_emul_geometry(int* tr, int* tc, int* apron) {
    *tr = ResGeom::TR; *tc = ResGeom::TC; *apron = ResGeom::APRON;
    return 0;
}
