// TEST INFRASTRUCTURE ONLY -- never linked into, loaded by or reachable from the package.
//
// The device code of the streaming kernels (kernels.cuh: step_stream and stepK_stream with their tile functions) run on
// the CPU through tests/host_emul/simt.h, one warp at a time.  tests/test_stream_emulation.py compares the result with
// the oracle, so a kernel edit can be pinned bitwise BEFORE it costs GPU time (and profiles/validated_sass.json tells
// whether an edit changed the shipped device code at all).
#define FDTD_HOST_EMUL 1
#include "simt.h"
#include "kernels.cuh"
#include "edge.cuh"

namespace fdtd { alignas(16) unsigned char ring_raw[256 * 1024]; }   // the dynamic shared memory of stepK_stream

using namespace fdtd;

extern "C" {

struct EmuSrc { int32_t kind, row, col, pad; double mag, magz; };

struct StreamScene {
    const double* a[3];         // generation A: h, ex, ey ((rows+1) x pitch each, whole grid)
    double* b[3];               // generation B
    const double* muc; const double* epc;
    int64_t pitch;
    int32_t NR, NC;
    int32_t absorb[4];
    double S, oms, last_mag;
    int32_t n_src; EmuSrc src[FDTD_MAX_SRC];
    int32_t n_rect;
    int32_t rh[FDTD_MAX_RECT][4], rex[FDTD_MAX_RECT][4], rey[FDTD_MAX_RECT][4];
};

static void fill(StepParams<double>& P, const StreamScene& sc) {
    std::memset(&P, 0, sizeof(P));
    P.h = sc.a[0]; P.ex = sc.a[1]; P.ey = sc.a[2]; P.hn = sc.b[0]; P.exn = sc.b[1]; P.eyn = sc.b[2];
    P.muc = sc.muc; P.epc = sc.epc; P.pitch = sc.pitch; P.NR = sc.NR; P.NC = sc.NC;
    P.row_off = 0; P.lrows = sc.NR + 1;
    P.absorb_u = sc.absorb[0]; P.absorb_d = sc.absorb[1]; P.absorb_l = sc.absorb[2]; P.absorb_r = sc.absorb[3];
    P.S = sc.S; P.oms = sc.oms; P.last_mag = sc.last_mag;
    P.n_src = sc.n_src;
    for (int s = 0; s < sc.n_src; ++s) {
        P.src[s].kind = sc.src[s].kind; P.src[s].row = sc.src[s].row; P.src[s].col = sc.src[s].col; P.src[s].pad = 0;
        P.src[s].mag = sc.src[s].mag; P.src[s].magz = sc.src[s].magz;
    }
    P.n_rect = sc.n_rect;
    for (int r = 0; r < sc.n_rect; ++r) {
        P.rh[r] = RectI{sc.rh[r][0], sc.rh[r][1], sc.rh[r][2], sc.rh[r][3]};
        P.rex[r] = RectI{sc.rex[r][0], sc.rex[r][1], sc.rex[r][2], sc.rex[r][3]};
        P.rey[r] = RectI{sc.rey[r][0], sc.rey[r][1], sc.rey[r][2], sc.rey[r][3]};
    }
}

// One launch of step_stream<double, V> over rows [row_lo, row_hi): the kernel body itself, grid and all.
int emul_step_stream(const StreamScene* sc, int V, int tile_rows, int warps, int row_lo, int row_hi) {
    StepParams<double> P;
    fill(P, *sc);
    P.row_lo = row_lo; P.row_hi = row_hi; P.tile_rows = tile_rows;
    const int W = 32 * V;
    const int gx = (P.NC + 1 + W * warps - 1) / (W * warps), gy = (row_hi - row_lo + tile_rows - 1) / tile_rows;
    simt::g_block_dim = simt::Idx{(unsigned)(warps * 32), 1, 1};
    simt::g_grid_dim = simt::Idx{(unsigned)gx, (unsigned)gy, 1};
    for (int by = 0; by < gy; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            simt::g_block_idx = simt::Idx{(unsigned)bx, (unsigned)by, 0};
            for (int w = 0; w < warps; ++w) {
                if (V == 2) simt::run_warp(w, [&] { step_stream<double, 2>(P); });
                else if (V == 4) simt::run_warp(w, [&] { step_stream<double, 4>(P); });
                else return -1;
            }
        }
    return 0;
}

// One launch of stepK_stream<double, 2, K>: K calls for every tile of rows [row_lo, row_hi) whose skip byte is 0.
// skip: (tiles_y x tiles_x) bytes, as fdtd_plan_frame returns them.
// variant 0 = stepK_stream (shipped), 1 = stepK_lean (unrolled row loop, running pointers).
int emul_stepK_stream(const StreamScene* sc, int K, int tile_rows, int warps, int row_lo, int row_hi,
                      const unsigned char* skip, int tiles_x, int tiles_y, int variant) {
    StepParams<double> P;
    fill(P, *sc);
    P.n_src = 0; P.n_rect = 0;                                   // as group_impl launches it
    P.row_lo = row_lo; P.row_hi = row_hi; P.tile_rows = tile_rows;
    const int gx = (tiles_x + warps - 1) / warps;
    simt::g_block_dim = simt::Idx{(unsigned)(warps * 32), 1, 1};
    simt::g_grid_dim = simt::Idx{(unsigned)gx, (unsigned)tiles_y, 1};
    for (int by = 0; by < tiles_y; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            simt::g_block_idx = simt::Idx{(unsigned)bx, (unsigned)by, 0};
            for (int w = 0; w < warps; ++w) {
                if (variant == 3) {                              // stepK_skew_umu: uniform plane + the stage chain cut
                    P.mu_uniform = sc->muc[0];
                    switch (K) {
                        case 2: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 2>(P, skip, tiles_x); }); break;
                        case 3: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 3>(P, skip, tiles_x); }); break;
                        case 4: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 4>(P, skip, tiles_x); }); break;
                        case 5: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 5>(P, skip, tiles_x); }); break;
                        case 6: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 6>(P, skip, tiles_x); }); break;
                        case 7: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 7>(P, skip, tiles_x); }); break;
                        case 8: simt::run_warp(w, [&] { stepK_skew_umu<double, 2, 8>(P, skip, tiles_x); }); break;
                        default: return -1;
                    }
                    continue;
                }
                if (variant == 2) {                              // stepK_lean_umu: the value of the (uniform) plane, as the host passes it
                    P.mu_uniform = sc->muc[0];
                    switch (K) {
                        case 2: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 2>(P, skip, tiles_x); }); break;
                        case 3: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 3>(P, skip, tiles_x); }); break;
                        case 4: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 4>(P, skip, tiles_x); }); break;
                        case 5: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 5>(P, skip, tiles_x); }); break;
                        case 6: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 6>(P, skip, tiles_x); }); break;
                        case 7: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 7>(P, skip, tiles_x); }); break;
                        case 8: simt::run_warp(w, [&] { stepK_lean_umu<double, 2, 8>(P, skip, tiles_x); }); break;
                        default: return -1;
                    }
                    continue;
                }
                if (variant == 1) {
                    switch (K) {
                        case 2: simt::run_warp(w, [&] { stepK_lean<double, 2, 2>(P, skip, tiles_x); }); break;
                        case 3: simt::run_warp(w, [&] { stepK_lean<double, 2, 3>(P, skip, tiles_x); }); break;
                        case 4: simt::run_warp(w, [&] { stepK_lean<double, 2, 4>(P, skip, tiles_x); }); break;
                        case 5: simt::run_warp(w, [&] { stepK_lean<double, 2, 5>(P, skip, tiles_x); }); break;
                        case 6: simt::run_warp(w, [&] { stepK_lean<double, 2, 6>(P, skip, tiles_x); }); break;
                        case 7: simt::run_warp(w, [&] { stepK_lean<double, 2, 7>(P, skip, tiles_x); }); break;
                        case 8: simt::run_warp(w, [&] { stepK_lean<double, 2, 8>(P, skip, tiles_x); }); break;
                        default: return -1;
                    }
                    continue;
                }
                switch (K) {
                    case 2: simt::run_warp(w, [&] { stepK_stream<double, 2, 2>(P, skip, tiles_x); }); break;
                    case 3: simt::run_warp(w, [&] { stepK_stream<double, 2, 3>(P, skip, tiles_x); }); break;
                    case 4: simt::run_warp(w, [&] { stepK_stream<double, 2, 4>(P, skip, tiles_x); }); break;
                    case 5: simt::run_warp(w, [&] { stepK_stream<double, 2, 5>(P, skip, tiles_x); }); break;
                    case 6: simt::run_warp(w, [&] { stepK_stream<double, 2, 6>(P, skip, tiles_x); }); break;
                    case 7: simt::run_warp(w, [&] { stepK_stream<double, 2, 7>(P, skip, tiles_x); }); break;
                    case 8: simt::run_warp(w, [&] { stepK_stream<double, 2, 8>(P, skip, tiles_x); }); break;
                    default: return -1;
                }
            }
        }
    return 0;
}

// One launch of stepK_lean_split<K> (kernels.cuh): the plain tiles of a K-step group under split field storage.  The
// scene's a[] / b[] point at split buffers (fp32 hi plane, then the bf16 lo plane split_lo_off bytes behind it).
int emul_stepK_lean_split(const StreamScene* sc, int K, int tile_rows, int warps, int row_lo, int row_hi,
                          const unsigned char* skip, int tiles_x, int tiles_y, long long split_lo_off) {
    StepParams<double> P;
    fill(P, *sc);
    P.n_src = 0; P.n_rect = 0;
    P.row_lo = row_lo; P.row_hi = row_hi; P.tile_rows = tile_rows;
    P.split_lo_off = split_lo_off;
    const int gx = (tiles_x + warps - 1) / warps;
    simt::g_block_dim = simt::Idx{(unsigned)(warps * 32), 1, 1};
    simt::g_grid_dim = simt::Idx{(unsigned)gx, (unsigned)tiles_y, 1};
    for (int by = 0; by < tiles_y; ++by)
        for (int bx = 0; bx < gx; ++bx) {
            simt::g_block_idx = simt::Idx{(unsigned)bx, (unsigned)by, 0};
            for (int w = 0; w < warps; ++w)
                switch (K) {
                    case 2: simt::run_warp(w, [&] { stepK_lean_split<2>(P, skip, tiles_x); }); break;
                    case 3: simt::run_warp(w, [&] { stepK_lean_split<3>(P, skip, tiles_x); }); break;
                    case 4: simt::run_warp(w, [&] { stepK_lean_split<4>(P, skip, tiles_x); }); break;
                    case 5: simt::run_warp(w, [&] { stepK_lean_split<5>(P, skip, tiles_x); }); break;
                    case 6: simt::run_warp(w, [&] { stepK_lean_split<6>(P, skip, tiles_x); }); break;
                    case 7: simt::run_warp(w, [&] { stepK_lean_split<7>(P, skip, tiles_x); }); break;
                    case 8: simt::run_warp(w, [&] { stepK_lean_split<8>(P, skip, tiles_x); }); break;
                    default: return -1;
                }
        }
    return 0;
}

}  // extern "C"

// split storage: every tile through the all-features instantiation with SPLIT = 1
template <int K>
static void edge_tile_split(const StepParams<double>& P, const EdgeTile* tl, int n_tiles, const ResSources<double>& R) {
    simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_ALL, 1>(P, tl, n_tiles, R, 0); });
}

// one warp = one edge tile, through the instantiation of its class (edge.cuh EDGE_CLASSES)
template <int K>
static int edge_tile_class(int cls, const StepParams<double>& P, const EdgeTile* tl, int n_tiles, const ResSources<double>& R) {
    switch (cls) {
        case EDGE_COLS: simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_COLS>(P, tl, n_tiles, R, 0); }); return 0;
        case EDGE_ROWS: simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_ROWS>(P, tl, n_tiles, R, 0); }); return 0;
        case EDGE_SRC: simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_SRC>(P, tl, n_tiles, R, 0); }); return 0;
        case EDGE_COLS | EDGE_SRC: simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_COLS | EDGE_SRC>(P, tl, n_tiles, R, 0); }); return 0;
        case EDGE_ALL: simt::run_warp(0, [&] { stepK_edge<double, 2, K, EDGE_ALL>(P, tl, n_tiles, R, 0); }); return 0;
        default: return -2;
    }
}

extern "C" {

// One launch of stepK_edge<double, 2, K>: K calls for every listed tile (tiles: n x {tx, ta, tb, class or 0 = all features}).  The pulse records
// of the K calls come as the device tables of resident mode do: count[K], last[K], src[K][stride].
int emul_stepK_edge(const StreamScene* sc, int K, const int32_t* tiles, int n_tiles, const int32_t* count, const double* last,
                    const EmuSrc* src, int stride, int n_probe, const int32_t* probe_ij, double* probe_out,
                    long long split_lo_off) {
    StepParams<double> P;
    fill(P, *sc);
    P.n_src = 0;                                                 // the kernel reads the records itself
    P.row_lo = 0; P.row_hi = P.NR + 1; P.tile_rows = 0;
    P.n_probe = n_probe;
    for (int k = 0; k < n_probe; ++k) { P.probe_i[k] = probe_ij[2 * k]; P.probe_j[k] = probe_ij[2 * k + 1]; }
    P.probe_out = probe_out;
    std::vector<SrcStep<double>> recs((size_t)K * std::max(stride, 1));
    for (int m = 0; m < K; ++m)
        for (int k = 0; k < stride; ++k) {
            const EmuSrc& e = src[(size_t)m * stride + k];
            SrcStep<double>& q = recs[(size_t)m * stride + k];
            q.kind = e.kind; q.row = e.row; q.col = e.col; q.pad = 0; q.mag = e.mag; q.magz = e.magz;
        }
    ResSources<double> R{stride ? count : nullptr, last, recs.data(), stride};
    simt::g_block_dim = simt::Idx{32, 1, 1};
    simt::g_grid_dim = simt::Idx{(unsigned)n_tiles, 1, 1};
    const EdgeTile* tl = reinterpret_cast<const EdgeTile*>(tiles);
    for (int b = 0; b < n_tiles; ++b) {
        simt::g_block_idx = simt::Idx{(unsigned)b, 0, 0};
        const int cls = tl[b].pad ? tl[b].pad : (int)EDGE_ALL;   // the instantiation the host launches this tile with
        int rc = -1;
        if (split_lo_off) {                                      // split storage (FDTD_F32X): the planes are hi / lo pairs
            P.split_lo_off = split_lo_off;
            switch (K) {
                case 1: edge_tile_split<1>(P, tl, n_tiles, R); break;
                case 2: edge_tile_split<2>(P, tl, n_tiles, R); break;
                case 3: edge_tile_split<3>(P, tl, n_tiles, R); break;
                case 4: edge_tile_split<4>(P, tl, n_tiles, R); break;
                case 5: edge_tile_split<5>(P, tl, n_tiles, R); break;
                case 6: edge_tile_split<6>(P, tl, n_tiles, R); break;
                case 7: edge_tile_split<7>(P, tl, n_tiles, R); break;
                case 8: edge_tile_split<8>(P, tl, n_tiles, R); break;
                default: return -1;
            }
            continue;
        }
        switch (K) {
            case 1: rc = edge_tile_class<1>(cls, P, tl, n_tiles, R); break;
            case 2: rc = edge_tile_class<2>(cls, P, tl, n_tiles, R); break;
            case 3: rc = edge_tile_class<3>(cls, P, tl, n_tiles, R); break;
            case 4: rc = edge_tile_class<4>(cls, P, tl, n_tiles, R); break;
            case 5: rc = edge_tile_class<5>(cls, P, tl, n_tiles, R); break;
            case 6: rc = edge_tile_class<6>(cls, P, tl, n_tiles, R); break;
            case 7: rc = edge_tile_class<7>(cls, P, tl, n_tiles, R); break;
            case 8: rc = edge_tile_class<8>(cls, P, tl, n_tiles, R); break;
        }
        if (rc) return rc;
    }
    return 0;
}

}  // extern "C"
