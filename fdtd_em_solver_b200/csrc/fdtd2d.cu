// libfdtd2d.so -- host side of the C ABI declared in include/fdtd2d.h.
// Replaces the reference's Solver.update / Solver.solve loops
// (/root/reference/solver.py:169-257, :270-337) with launches of the fused
// sm_100a kernels in kernels.cuh.  No CPU fallback exists: every entry point
// that computes needs a CUDA device.
#include "fdtd2d.h"
#include "launch.h"

#include <dlfcn.h>
#include <nvtx3/nvToolsExt.h>

#include <algorithm>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

using fdtd::RectI;
using fdtd::StepParams;

namespace {

// NVTX ranges (the reference has no tracing; these show up in nsys / ncu --nvtx timelines).  Header-only
// NVTX3: a no-op costing one branch per call unless a tool injected itself into the process.
struct Range {
    explicit Range(const char* name) { nvtxRangePushA(name); }
    ~Range() { nvtxRangePop(); }
    Range(const Range&) = delete;
    Range& operator=(const Range&) = delete;
};

std::string g_create_error;

// ---- NCCL, bound at run time so that single-GPU use never needs it -------------------------
typedef struct ncclComm* ncclComm_t;
struct NcclId { char internal[128]; };
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(NcclId*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*Send)(const void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*Recv)(void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*GroupStart)() = nullptr;
    int (*GroupEnd)() = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool load(std::string& err) {
        if (lib) return true;
        const char* names[] = {"libnccl.so.2", "libnccl.so"};
        for (const char* n : names) { lib = dlopen(n, RTLD_NOW | RTLD_GLOBAL); if (lib) break; }
        if (!lib) { err = std::string("dlopen(libnccl.so.2) failed: ") + dlerror(); return false; }
#define FDTD_SYM(field, sym) field = reinterpret_cast<decltype(field)>(dlsym(lib, sym)); \
        if (!field) { err = std::string("missing NCCL symbol ") + sym; lib = nullptr; return false; }
        FDTD_SYM(GetUniqueId, "ncclGetUniqueId") FDTD_SYM(CommInitRank, "ncclCommInitRank")
        FDTD_SYM(CommDestroy, "ncclCommDestroy") FDTD_SYM(Send, "ncclSend") FDTD_SYM(Recv, "ncclRecv")
        FDTD_SYM(GroupStart, "ncclGroupStart") FDTD_SYM(GroupEnd, "ncclGroupEnd")
        FDTD_SYM(GetErrorString, "ncclGetErrorString")
#undef FDTD_SYM
        return true;
    }
} g_nccl;

}  // namespace

struct FrameMaps {                   // per group size K: which tiles the K-step kernel skips and the frame passes cover
    std::vector<uint8_t> key;       // active-source signature + geometry the maps were built for
    std::vector<uint8_t> h_skip, h_mask[8];      // h_mask[d]: frame tiles dilated by d rings (d = 0 is the frame itself)
    unsigned char* d_skip = nullptr; unsigned char* d_mask[8] = {};
    std::vector<int2> h_list[8];                 // compact (block column, row tile) launch lists of the frame passes
    int2* d_list[8] = {};
    std::vector<fdtd::EdgeTile> h_edge;          // frame mode 3: the tiles of stepK_edge (edge.cuh), those next to a slab
    fdtd::EdgeTile* d_edge = nullptr;            // interface first (n_edge_near of them)
    size_t d_edge_cap = 0;
    int n_edge_near = 0;
    struct EdgeRun { int first, count, flags; };  // consecutive tiles of one class (the kernel instantiation `flags`)
    std::vector<EdgeRun> edge_runs_near, edge_runs_far;
    // fdtd_run_streamed: the same tiles grouped by row band (band = first row / band_rows), a second device copy of the
    // edge tiles sorted by (band, class)
    int64_t band_rows = 0;
    std::vector<int> band_ty0, band_ny;
    std::vector<std::vector<EdgeRun>> band_runs;
    fdtd::EdgeTile* d_band = nullptr;
    size_t d_band_cap = 0;
    size_t d_list_cap = 0, d_skip_cap = 0, d_mask_cap = 0;
    int nty = 0, ntx2 = 0, ntx1 = 0;     // K-step tile grid (nty x ntx2) and frame warp tiles per row (ntx1)
    std::vector<int64_t> row_cells;      // cells the plain K-step kernel advances per tile row (measurement hook)
    int frame_tile_rows = 0, frame_vec = 0, frame_nty = 0;   // the frame passes use smaller tiles than the K-step kernel
    int64_t k_row_lo = 0, k_row_hi = 0;                      // rows the K-step tiles own
    void release() {
        if (d_skip) cudaFree(d_skip);
        if (d_edge) cudaFree(d_edge);
        if (d_band) cudaFree(d_band);
        for (int d = 0; d < 8; ++d) { if (d_mask[d]) cudaFree(d_mask[d]); if (d_list[d]) cudaFree(d_list[d]); }
    }
};

struct fdtd_ctx {
    fdtd_config cfg{};
    int64_t pitch = 0, lrows = 0, halo_top = 0, halo_bot = 0, row_off = 0, owned = 0, host_row0 = 0;
    int halo_depth = 1;             // rows of h/ex/ey kept from each neighbouring slab: 1 = one row above + ex[row_end]
                                    // (exchange every step); D >= 2 = D rows above AND below (deep halos: ONE exchange
                                    // per temporal-blocking group of up to D-1 steps, fdtd_set_halo_depth)
    size_t esize = 8;
    void* gen[4][3] = {};           // [generation][h,ex,ey]: 2 for ping-pong, +1 / +2 scratch for temporal blocking
    int ngen = 2;                   // generations bound
    int temporal = 1;               // leapfrog steps per HBM pass in fdtd_run (1..8)
    bool auto_tile_rows = false, auto_warps = false;   // launch geometry left to the library (follows `temporal`)
    FrameMaps fm[9];                // indexed by group size K (2..8)
    void* epc = nullptr; void* muc = nullptr;
    bool bound = false;
    int cur = 0;

    int n_src = 0;
    fdtd_source src[64];
    int64_t table_len = 0;
    std::vector<double> mag, magz, last;
    std::vector<uint8_t> active;

    int n_rect = 0;
    RectI rh[FDTD_MAX_RECT], rex[FDTD_MAX_RECT], rey[FDTD_MAX_RECT];

    int n_probe = 0;
    int probe_i[FDTD_MAX_PROBE], probe_j[FDTD_MAX_PROBE];
    double* probe_dev = nullptr;
    int64_t probe_cap = 0, probe_count = 0;

    cudaStream_t stream = nullptr, copy_stream = nullptr, comm_stream = nullptr, edge_stream = nullptr;
    cudaEvent_t ev_t0 = nullptr, ev_t1 = nullptr, ev_bnd = nullptr, ev_comm = nullptr, ev_int = nullptr, ev_p3 = nullptr;
    cudaEvent_t ev_packed[2] = {nullptr, nullptr}, ev_free[2] = {nullptr, nullptr};
    double* snap_dev[2] = {nullptr, nullptr};
    double* snap_pin[2] = {nullptr, nullptr};        // pinned host ring: snapshots reach pageable user memory through it
    struct SnapCopy { double* dst; const double* src; size_t bytes; };
    // realtime view (fdtd_view_begin / fdtd_view_end): a ring of two decimated pictures of h in flight
    struct View { double* dev = nullptr; double* pin = nullptr; size_t cap = 0; cudaEvent_t done = nullptr;
                  int64_t rows = 0, cols = 0; bool pending = false; } view[2];
    int view_head = 0, view_tail = 0;                // next slot to fill / oldest pending slot
    // fdtd_run_streamed: device-to-host copies on their own stream (PCIe is full duplex), one event per row band
    cudaStream_t d2h_stream = nullptr, edge_stream2 = nullptr;
    std::vector<cudaEvent_t> band_up, band_done, band_edge;
    void* stage = nullptr; size_t stage_bytes = 0;

    ncclComm_t comm = nullptr;
    int rank = 0, nranks = 1;

    // resident mode (step_resident: many steps per cooperative launch, for grids that live in L2)
    int resident = 0;
    int64_t resident_max_cells = 512 * 512;      // (rows+1)*(cols+1) up to which fdtd_run uses it
    void* rs_count = nullptr; void* rs_last = nullptr; void* rs_src = nullptr;   // per-call pulse records on the device
    bool rs_valid = false;
    int rs_stride = 0, res_blocks_max = 0;

    // how the non-plain tiles of a temporal-blocking group advance: 3 = ONE stepK_edge launch next to the plain kernel
    // (default since its B200 runs, profiles/r2_sweep_*), 0 = K masked single-step passes through scratch generations
    int frame_mode = 3;
    int kstep_variant = 1;           // 0 = stepK_stream, 1 = stepK_lean (same bits, fewer issue slots; default since round 2)
    int edge_rows = 32;              // rows per stepK_edge tile (experiment hook: FDTD_EDGE_ROWS)
    // split field storage (dtype FDTD_F32X): every tile is a stepK_edge tile; one tile list per group size
    // uniform mu_coef plane (mu_r = 1 everywhere: every script of the reference, the C5 recipe): found by a device scan
    // whenever the coefficients arrive; the plain K-step kernel then reads the value from its parameters (stepK_lean_umu)
    bool mu_uniform = false, mu_uniform_enable = true, mu_scanned = false;
    bool skew = true;                    // stepK_skew_umu (the stage chain cut: 788 vs 769 Gcell/s) for a uniform plane; FDTD_SKEW=0: stepK_lean_umu
    double mu_uniform_value = 0.0;
    void* scan_dev = nullptr;            // {T value (8 bytes), int differs}
    bool split = false;
    bool split_all_edge = false;         // FDTD_SPLIT_ALL_EDGE=1: every tile through stepK_edge (the first version of the mode)
    struct SplitTiles { fdtd::EdgeTile* dev = nullptr; int n = 0; } split_tiles[9];
    bool maps_pinned = false;        // fdtd_run_streamed: the frame maps in place serve every group of the run
    bool edge_class_force = false;   // measurement hook FDTD_EDGE_GENERIC=1: every edge tile through the all-features instantiation

    // optional: CUDA events around every launch of the plain K-step kernel on its own stream (bench.py: the roofline
    // line wants the dominant kernel's own duration, measured live in the timed region)
    bool ktime_on = false;
    std::vector<cudaEvent_t> kt_ev;
    int kt_n = 0;
    int64_t kt_cell_steps = 0;       // cell-updates of the timed launches

    fdtd_stats stats{};
    std::string err;
};

namespace {

int fail(fdtd_ctx* c, int code, const std::string& msg) {
    if (c) c->err = msg; else g_create_error = msg;
    return code;
}

#define CK(call) do { cudaError_t e_ = (call); if (e_ != cudaSuccess) \
    return fail(c, FDTD_ERR_CUDA, std::string(#call) + ": " + cudaGetErrorString(e_)); } while (0)
#define NCK(call) do { int r_ = (call); if (r_ != 0) \
    return fail(c, FDTD_ERR_COMM, std::string(#call) + ": " + g_nccl.GetErrorString(r_)); } while (0)

int64_t norm_index(int64_t v, int64_t len) {        // python slice.indices(), step 1
    if (v < 0) { v += len; if (v < 0) v = 0; }
    if (v > len) v = len;
    return v;
}

RectI norm_rect(const int64_t* r, int64_t nrows, int64_t ncols) {
    RectI o;
    o.r0 = (int)norm_index(r[0], nrows); o.r1 = (int)norm_index(r[1], nrows);
    o.c0 = (int)norm_index(r[2], ncols); o.c1 = (int)norm_index(r[3], ncols);
    return o;
}

template <typename T>
T* plane(void* base, fdtd_ctx* c, int64_t global_row) {
    return reinterpret_cast<T*>(base) + (global_row - c->row_off) * c->pitch;
}

int ensure_stage(fdtd_ctx* c, size_t bytes) {
    if (c->stage_bytes >= bytes) return 0;
    if (c->stage) cudaFree(c->stage);
    c->stage = nullptr; c->stage_bytes = 0;
    CK(cudaMalloc(&c->stage, bytes));
    c->stage_bytes = bytes;
    return 0;
}

// global host rows [g0, g0+n) x ncols (ld host_ld) -> device plane `base`
template <typename T>
int h2d_plane(fdtd_ctx* c, void* base, const double* host, int64_t host_ld, int64_t g0, int64_t n, int64_t ncols) {
    if (n <= 0) return 0;
    T* dst = plane<T>(base, c, g0);
    if (!host) { CK(cudaMemset2DAsync(dst, c->pitch * sizeof(T), 0, ncols * sizeof(T), n, c->stream)); return 0; }
    const double* src = host + (g0 - c->host_row0) * host_ld;
    if (sizeof(T) == 8) {
        CK(cudaMemcpy2DAsync(dst, c->pitch * 8, src, host_ld * 8, ncols * 8, n, cudaMemcpyHostToDevice, c->stream));
        return 0;
    }
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(32 << 20) / (ncols * 8));
    if (int rc = ensure_stage(c, (size_t)std::min(chunk, n) * ncols * 8)) return rc;
    for (int64_t a = 0; a < n; a += chunk) {
        const int64_t m = std::min(chunk, n - a);
        CK(cudaMemcpy2DAsync(c->stage, ncols * 8, src + a * host_ld, host_ld * 8, ncols * 8, m,
                             cudaMemcpyHostToDevice, c->stream));
        dim3 grid((unsigned)((ncols + 255) / 256), (unsigned)std::min<int64_t>(m, 65535));
        CK(fdtd::launch_plane_from_f64<T>(grid, c->stream, dst + a * c->pitch, c->pitch, (const double*)c->stage, ncols,
                                          (int)m, (int)ncols));
        c->stats.kernel_launches++;
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

template <typename T>
int d2h_plane(fdtd_ctx* c, const void* base, double* host, int64_t host_ld, int64_t g0, int64_t n, int64_t ncols) {
    if (n <= 0 || !host) return 0;
    const T* src = plane<T>(const_cast<void*>(base), c, g0);
    double* dst = host + (g0 - c->host_row0) * host_ld;
    if (sizeof(T) == 8) {
        CK(cudaMemcpy2DAsync(dst, host_ld * 8, src, c->pitch * 8, ncols * 8, n, cudaMemcpyDeviceToHost, c->stream));
        return 0;
    }
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(32 << 20) / (ncols * 8));
    if (int rc = ensure_stage(c, (size_t)std::min(chunk, n) * ncols * 8)) return rc;
    fdtd::PatchList none; none.n = 0;
    for (int64_t a = 0; a < n; a += chunk) {
        const int64_t m = std::min(chunk, n - a);
        dim3 block(64, 4), grid((unsigned)((ncols + 63) / 64), (unsigned)std::min<int64_t>((m + 3) / 4, 65535));
        CK(fdtd::launch_plane_to_f64<T>(grid, block, c->stream, (double*)c->stage, ncols, src + a * c->pitch, c->pitch,
                                        (int)m, (int)ncols, 0, none));
        c->stats.kernel_launches++;
        CK(cudaMemcpy2DAsync(dst + a * host_ld, host_ld * 8, c->stage, ncols * 8, ncols * 8, m,
                             cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int next_gen(const fdtd_ctx* c, int g) { return (g + 1) % c->ngen; }

template <typename T>
int fill_params(fdtd_ctx* c, StepParams<T>& P, int64_t step, int a, int b) {
    P.h = (const T*)c->gen[a][0]; P.ex = (const T*)c->gen[a][1]; P.ey = (const T*)c->gen[a][2];
    P.muc = (const T*)c->muc; P.epc = (const T*)c->epc;
    P.hn = (T*)c->gen[b][0]; P.exn = (T*)c->gen[b][1]; P.eyn = (T*)c->gen[b][2];
    P.pitch = c->pitch;
    P.NR = (int)c->cfg.rows; P.NC = (int)c->cfg.cols;
    P.row_off = (int)c->row_off; P.lrows = (int)c->lrows;
    P.tile_rows = c->cfg.tile_rows;
    P.absorb_u = !c->cfg.reflect[0]; P.absorb_d = !c->cfg.reflect[1];
    P.absorb_l = !c->cfg.reflect[2]; P.absorb_r = !c->cfg.reflect[3];
    P.S = (T)c->cfg.courant; P.oms = (T)c->cfg.one_minus_courant;
    P.n_src = 0; P.last_mag = T(0);
    if (c->n_src > 0) {
        if (step < 0 || step >= c->table_len)
            return fail(c, FDTD_ERR_RANGE, "step " + std::to_string(step) + " outside the source tables (length " +
                                               std::to_string(c->table_len) + ")");
        for (int p = 0; p < c->n_src; ++p) {
            if (!c->active[(size_t)p * c->table_len + step]) continue;
            if (P.n_src >= FDTD_MAX_SRC) return fail(c, FDTD_ERR_ARG, "more than 16 pulses active in one call");
            fdtd::SrcStep<T>& q = P.src[P.n_src++];
            q.kind = c->src[p].kind; q.row = (int)c->src[p].row; q.col = (int)c->src[p].col; q.pad = 0;
            q.mag = (T)c->mag[(size_t)p * c->table_len + step];
            q.magz = (T)c->magz[(size_t)p * c->table_len + step];
        }
        P.last_mag = (T)c->last[step];
    }
    P.n_rect = c->n_rect;
    for (int r = 0; r < c->n_rect; ++r) { P.rh[r] = c->rh[r]; P.rex[r] = c->rex[r]; P.rey[r] = c->rey[r]; }
    P.n_probe = c->n_probe;
    for (int k = 0; k < c->n_probe; ++k) { P.probe_i[k] = c->probe_i[k]; P.probe_j[k] = c->probe_j[k]; }
    P.probe_out = nullptr;
    P.tile_mask = nullptr; P.mask_stride = 0; P.tile_list = nullptr;
    P.split_lo_off = c->split ? (long long)c->lrows * c->pitch * 4 : 0;
    P.mu_uniform = (T)c->mu_uniform_value;
    return 0;
}

// Does every stored cell of the mu_coef plane hold one value?  (Rows: every h-row this context stores, halos included;
// columns 0 .. NC-1 -- what any plain K-step tile can load.)  One pass over the plane, about a millisecond for 4 GB.
template <typename T>
int scan_mu_uniform_t(fdtd_ctx* c) {
    c->mu_uniform = false;
    c->mu_scanned = true;
    if (!c->mu_uniform_enable || !c->muc) return 0;
    const int64_t rows = std::min<int64_t>(c->cfg.row_end + c->halo_bot, c->cfg.rows) - c->row_off;
    if (rows < 1) return 0;
    if (!c->scan_dev) CK(cudaMalloc(&c->scan_dev, 16));
    CK(cudaMemsetAsync(c->scan_dev, 0, 16, c->stream));
    dim3 grid((unsigned)((c->cfg.cols + 255) / 256), (unsigned)std::min<int64_t>(rows, 4096));
    CK(fdtd::launch_plane_uniform<T>(grid, c->stream, (const T*)c->muc, c->pitch, (int)rows, (int)c->cfg.cols,
                                     (T*)c->scan_dev, (int*)((char*)c->scan_dev + 8)));
    c->stats.kernel_launches++;
    unsigned char host[16];
    CK(cudaMemcpyAsync(host, c->scan_dev, 16, cudaMemcpyDeviceToHost, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    int differs; T value;
    std::memcpy(&value, host, sizeof(T)); std::memcpy(&differs, host + 8, 4);
    c->mu_uniform = differs == 0;
    c->mu_uniform_value = (double)value;
    return 0;
}
int scan_mu_uniform(fdtd_ctx* c) {
    return c->cfg.dtype != FDTD_F32 ? scan_mu_uniform_t<double>(c) : scan_mu_uniform_t<float>(c);
}

template <typename T>
cudaError_t launch_stream(const StepParams<T>& P, int V, int WARPS, cudaStream_t s, int n_list) {
    const int ncol_tiles = (P.NC + 1 + 32 * V * WARPS - 1) / (32 * V * WARPS);
    const int nrow_tiles = (P.row_hi - P.row_lo + P.tile_rows - 1) / P.tile_rows;
    dim3 grid((unsigned)ncol_tiles, (unsigned)nrow_tiles);
    if (P.tile_list) grid = dim3((unsigned)n_list, 1);
    return fdtd::launch_step_stream<T>(P, V, WARPS, grid, s);
}

template <typename T>
int launch_rows(fdtd_ctx* c, StepParams<T>& P, int64_t row_lo, int64_t row_hi, cudaStream_t s, int n_list = 0,
                int vec_override = 0) {
    if (row_hi <= row_lo) return 0;
    if (P.tile_list && n_list == 0) return 0;
    P.row_lo = (int)row_lo; P.row_hi = (int)row_hi;
    if (c->cfg.kernel == FDTD_KERNEL_EXACT) {
        dim3 grid((unsigned)((P.NC + 1 + 255) / 256), (unsigned)std::min<int64_t>(row_hi - row_lo, 65535));
        CK(fdtd::launch_step_exact<T>(P, grid, s));
    } else {
        const int v = vec_override ? vec_override : c->cfg.vec, w = c->cfg.block_warps;
        constexpr int VA = sizeof(T) == 8 ? 2 : 4, VB = 2 * VA;
        if (w < 1 || w > 8) return fail(c, FDTD_ERR_ARG, "block_warps must be 1..8");
        if (v != VA && v != VB) return fail(c, FDTD_ERR_ARG, "unsupported vec (2/4 for f64, 4/8 for f32)");
        CK(launch_stream<T>(P, v, w, s, n_list));
    }
    c->stats.kernel_launches++;
    return 0;
}

int ensure_probe_capacity(fdtd_ctx* c, int64_t extra_steps) {
    if (c->n_probe == 0) return 0;
    const int64_t need = c->probe_count + extra_steps;
    if (need <= c->probe_cap) return 0;
    const int64_t cap = std::max<int64_t>(need, c->probe_cap * 2);
    double* nb = nullptr;
    CK(cudaMalloc(&nb, (size_t)cap * c->n_probe * 3 * sizeof(double)));
    CK(cudaMemsetAsync(nb, 0, (size_t)cap * c->n_probe * 3 * sizeof(double), c->stream));
    if (c->probe_dev) {
        CK(cudaMemcpyAsync(nb, c->probe_dev, (size_t)c->probe_count * c->n_probe * 3 * sizeof(double),
                           cudaMemcpyDeviceToDevice, c->stream));
        CK(cudaStreamSynchronize(c->stream));
        cudaFree(c->probe_dev);
    }
    c->probe_dev = nb; c->probe_cap = cap;
    return 0;
}

int nccl_dtype(fdtd_ctx* c) { return c->cfg.dtype == FDTD_F64 ? 8 : 7; }

// Deep halos: refresh the D rows of h/ex/ey stored from each neighbour with that neighbour's D outermost owned rows
// of generation g -- one contiguous block of D x pitch elements per array and direction.
int exchange_deep(fdtd_ctx* c, int g, cudaStream_t s) {
    Range nvtx("fdtd:halo_exchange_deep");
    const int64_t D = c->halo_depth;
    const size_t es = c->esize, count = (size_t)(D * c->pitch);
    auto row = [&](int f, int64_t grow) { return (char*)c->gen[g][f] + (size_t)((grow - c->row_off) * c->pitch) * es; };
    const bool has_up = c->rank > 0, has_down = c->rank + 1 < c->nranks;
    const int dt = nccl_dtype(c);
    NCK(g_nccl.GroupStart());
    for (int f = 0; f < 3; ++f) {
        if (has_down) {
            NCK(g_nccl.Send(row(f, c->cfg.row_end - D), count, dt, c->rank + 1, c->comm, s));
            NCK(g_nccl.Recv(row(f, c->cfg.row_end), count, dt, c->rank + 1, c->comm, s));
        }
        if (has_up) {
            NCK(g_nccl.Send(row(f, c->cfg.row_begin), count, dt, c->rank - 1, c->comm, s));
            NCK(g_nccl.Recv(row(f, c->cfg.row_begin - D), count, dt, c->rank - 1, c->comm, s));
        }
    }
    NCK(g_nccl.GroupEnd());
    c->stats.halo_bytes_sent += (int64_t)((has_down ? 3 : 0) + (has_up ? 3 : 0)) * (int64_t)(D * c->cfg.cols * es);
    return 0;
}

// Exchange the halo rows of generation g on `s` (3 rows down, 1 row up).
int exchange(fdtd_ctx* c, int g, cudaStream_t s) {
    if (!c->comm) return 0;
    if (c->halo_depth > 1) return exchange_deep(c, g, s);
    Range nvtx("fdtd:halo_exchange");
    const int64_t NC = c->cfg.cols;
    const size_t es = c->esize;
    auto row = [&](int f, int64_t grow) { return (char*)c->gen[g][f] + (size_t)((grow - c->row_off) * c->pitch) * es; };
    const bool has_up = c->rank > 0, has_down = c->rank + 1 < c->nranks;
    const int dt = nccl_dtype(c);
    NCK(g_nccl.GroupStart());
    if (has_down) {
        NCK(g_nccl.Send(row(0, c->cfg.row_end - 1), (size_t)NC, dt, c->rank + 1, c->comm, s));
        NCK(g_nccl.Send(row(1, c->cfg.row_end - 1), (size_t)NC, dt, c->rank + 1, c->comm, s));
        NCK(g_nccl.Send(row(2, c->cfg.row_end - 1), (size_t)NC + 1, dt, c->rank + 1, c->comm, s));
        NCK(g_nccl.Recv(row(1, c->cfg.row_end), (size_t)NC, dt, c->rank + 1, c->comm, s));
        c->stats.halo_bytes_sent += (int64_t)((3 * NC + 1) * es);
    }
    if (has_up) {
        NCK(g_nccl.Send(row(1, c->cfg.row_begin), (size_t)NC, dt, c->rank - 1, c->comm, s));
        NCK(g_nccl.Recv(row(0, c->cfg.row_begin - 1), (size_t)NC, dt, c->rank - 1, c->comm, s));
        NCK(g_nccl.Recv(row(1, c->cfg.row_begin - 1), (size_t)NC, dt, c->rank - 1, c->comm, s));
        NCK(g_nccl.Recv(row(2, c->cfg.row_begin - 1), (size_t)NC + 1, dt, c->rank - 1, c->comm, s));
        c->stats.halo_bytes_sent += (int64_t)(NC * es);
    }
    NCK(g_nccl.GroupEnd());
    return 0;
}

template <typename T>
int step_impl(fdtd_ctx* c, int64_t step) {
    Range nvtx("fdtd:single_step");
    StepParams<T> P;
    const int dst = next_gen(c, c->cur);
    if (int rc = fill_params<T>(c, P, step, c->cur, dst)) return rc;
    if (c->n_probe) {
        if (int rc = ensure_probe_capacity(c, 1)) return rc;
        P.probe_out = c->probe_dev + (size_t)c->probe_count * c->n_probe * 3;
    }
    const int64_t lo = c->cfg.row_begin;
    const int64_t hi = c->cfg.row_end + (c->cfg.row_end == c->cfg.rows ? 1 : 0);   // last slab also owns ex[rows]
    if (c->auto_tile_rows) P.tile_rows = 6;        // cfg.tile_rows follows the K-step kernel; a lone single step wants short tiles
    if (!c->comm) {
        if (int rc = launch_rows<T>(c, P, lo, hi, c->stream)) return rc;
    } else if (c->halo_depth > 1) {
        // deep halos: single steps are the exception here (remainders, snapshot cadence), so no strip overlap -- all
        // owned rows in one launch once the previous exchange has landed, then the D-row blocks travel
        CK(cudaStreamWaitEvent(c->stream, c->ev_comm, 0));
        if (int rc = launch_rows<T>(c, P, lo, hi, c->stream)) return rc;
        CK(cudaEventRecord(c->ev_p3, c->stream));
        CK(cudaStreamWaitEvent(c->comm_stream, c->ev_p3, 0));
        if (int rc = exchange(c, dst, c->comm_stream)) return rc;
        CK(cudaEventRecord(c->ev_comm, c->comm_stream));
        CK(cudaEventRecord(c->ev_int, c->stream));
    } else {
        // Interface strips on a high-priority stream, their rows sent over NCCL on the comm stream, the
        // interior on the main stream -- all three overlap.  Dependencies of step n (A -> B):
        //   edge(n)     after interior(n-1) [it overwrites A strip rows interior(n-1) still reads] and comm(n-1)
        //   comm(n)     after edge(n)
        //   interior(n) after edge(n-1) [it reads generation A strip rows that edge(n-1) wrote]; same stream as interior(n-1)
        const int64_t tr = P.tile_rows;
        const bool has_up = c->rank > 0, has_down = c->rank + 1 < c->nranks;
        int64_t a = has_up ? std::min(lo + tr, hi) : lo;
        int64_t b = has_down ? std::max(a, c->cfg.row_end - tr) : hi;
        CK(cudaStreamWaitEvent(c->edge_stream, c->ev_int, 0));
        CK(cudaStreamWaitEvent(c->edge_stream, c->ev_comm, 0));
        if (int rc = launch_rows<T>(c, P, lo, a, c->edge_stream)) return rc;
        if (int rc = launch_rows<T>(c, P, b, hi, c->edge_stream)) return rc;
        CK(cudaEventRecord(c->ev_bnd, c->edge_stream));
        CK(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
        if (int rc = exchange(c, dst, c->comm_stream)) return rc;
        CK(cudaEventRecord(c->ev_comm, c->comm_stream));
        if (int rc = launch_rows<T>(c, P, a, b, c->stream)) return rc;
        CK(cudaEventRecord(c->ev_int, c->stream));
        CK(cudaStreamWaitEvent(c->stream, c->ev_bnd, 0));      // main-stream order now implies "all owned rows written"
    }
    c->cur = dst;
    if (c->n_probe) c->probe_count++;
    c->stats.steps++;
    c->stats.cell_updates += c->owned * c->cfg.cols;
    return 0;
}

// ---- temporal blocking (K steps per HBM pass) ----------------------------------------------------------
template <typename T> constexpr int kvec() { return sizeof(T) == 8 ? 2 : 4; }     // columns per lane of stepK_stream

// Frame mode 3: can stepK_edge (edge.cuh) take every non-plain tile of a K-step group of this context?  It handles all
// walls, pulse kinds, reflectors and probes; what it needs is the whole grid on this GPU or, on a y-slab, halos deep
// enough for the group (K+1 rows), and no TF/SF "right" line at column 0 (python's h[:, col-1] wraps to the far column).
bool edge_usable(const fdtd_ctx* c, int K) {
    if (c->frame_mode != 3 || c->n_src > FDTD_MAX_SRC || K < 2 || K > 8) return false;
    const bool whole = c->cfg.row_begin == 0 && c->cfg.row_end == c->cfg.rows;
    if (!whole && c->halo_depth < K + 1) return false;
    for (int p = 0; p < c->n_src; ++p)
        if (c->src[p].kind == FDTD_SRC_RIGHT && c->src[p].col == 0) return false;
    return true;
}

// Which K-step tiles must NOT be advanced by stepK_stream, and which single-step tiles advance them instead.
// Pure host logic (no CUDA): also exported as fdtd_plan_frame so that it can be tested without a GPU.
struct FrameBox { int64_t r0, r1, c0, c1; int what = fdtd::EDGE_SRC | fdtd::EDGE_MISC; };   // inclusive cell box of a wall-less "modifier" (+ what it is, for the edge classes)
struct FrameGeom { int64_t rows, cols, row_begin, row_end; int tile_rows, block_warps, K, V2;
                   int64_t halo_top = 0, halo_bot = 0; bool edge = false; int edge_rows = 32; };   // rows of h/ex/ey STORED beyond the owned ones (deep halos; 0 = classic)

void plan_frame(const FrameGeom& g, const std::vector<FrameBox>& mods, FrameMaps& m) {
    const int K = g.K, V2 = g.V2;
    const int MARGIN = (K + V2 - 1) / V2, OUT = (32 - 2 * MARGIN) * V2, LD = 32 * V2;
    const int64_t NR = g.rows, NC = g.cols;
    const int64_t lo = g.row_begin, hi = g.row_end + (g.row_end == NR ? 1 : 0);
    const int TR = g.tile_rows;                                    // rows per K-step tile
    // The frame passes run latency-bound "mixed" tiles, so they get the narrowest lanes and short tiles: more, shorter
    // CTAs and a thin dilation ring (profiles/sweep_r1n..p) instead of the K-step geometry.
    const int VF = V2, W1 = 32 * VF;
    int TRF = TR;                                                  // largest divisor of TR not above 8
    for (int d = 8; d >= 2; --d) if (TR % d == 0) { TRF = d; break; }
    if (TR <= 8) TRF = (TR % 2 == 0 && TR >= 8) ? TR / 2 : TR;
    // K-step tiles own rows [loK, hiK): they start K rows (rounded up to a frame tile) below the slab top and stop K
    // rows above its bottom, so that only a thin band of rows at the top/bottom (or at a slab interface) is frame.
    // Deep halos store K rows (and one more) of the neighbouring slab, so at an INTERFACE the K-step tiles reach the
    // slab's first / last owned row; at a real wall of the grid nothing changes.
    const bool deep_top = g.row_begin > 0 && g.halo_top >= K, deep_bot = g.row_end < NR && g.halo_bot >= K + 1;
    const int64_t loK = deep_top ? lo : lo + ((K + TRF - 1) / TRF) * (int64_t)TRF;
    int64_t hiK = deep_bot ? g.row_end : std::min<int64_t>(g.row_end, NR) - K;   // tb+K-1 <= last stored row - 1 and tb+K-2 <= NR-2
    hiK = hiK > lo ? lo + ((hiK - lo) / TRF) * TRF : lo;
    const int nty = hiK > loK ? (int)((hiK - loK + TR - 1) / TR) : 0, ntyF = (int)((hi - lo + TRF - 1) / TRF);
    const int ntx2 = (int)((NC + 1 + OUT - 1) / OUT), ntx1 = (int)((NC + 1 + W1 - 1) / W1);
    m.nty = nty; m.ntx2 = ntx2; m.ntx1 = ntx1; m.frame_tile_rows = TRF; m.frame_vec = VF; m.frame_nty = ntyF;
    m.k_row_lo = loK; m.k_row_hi = hiK;
    m.h_skip.assign((size_t)std::max(nty, 1) * ntx2, 0);
    // frame tiles: everything that is not completely covered by K-step tiles the K-step kernel really advances
    for (int d = 0; d < 8; ++d) { m.h_mask[d].assign(d < K ? (size_t)ntyF * ntx1 : 0, d == 0 ? 1 : 0); m.h_list[d].clear(); }
    std::vector<uint8_t> plain(ntx2);
    for (int ty = 0; ty < nty; ++ty) {
        const int64_t ta = loK + (int64_t)ty * TR, tb = std::min<int64_t>(ta + TR, hiK);
        // every row a stage touches (ta-K .. tb+K-1) is an OWNED plain interior row of this slab (by construction)
        const bool rows_ok = ta - K >= std::max<int64_t>(g.row_begin - g.halo_top, 0) && ta - (K - 1) >= 1 &&
                             tb + K - 2 <= NR - 2 && tb + K - 1 <= g.row_end + g.halo_bot - 1;
        for (int tx = 0; tx < ntx2; ++tx) {
            const int64_t jw = (int64_t)tx * OUT - (int64_t)MARGIN * V2;
            bool mixed = !rows_ok || jw < 1 || jw + LD - 1 > NC - 2;
            if (!mixed)
                for (const FrameBox& q : mods)
                    if (q.r0 <= tb + K && q.r1 >= ta - K - 1 && q.c0 <= jw + LD && q.c1 >= jw - 1) { mixed = true; break; }
            m.h_skip[(size_t)ty * ntx2 + tx] = mixed;
            plain[tx] = !mixed;
        }
        for (int64_t fy = (ta - lo) / TRF; fy < (tb - lo) / TRF && fy < ntyF; ++fy)       // ta, tb are frame-tile aligned
            for (int t1 = 0; t1 < ntx1; ++t1) {
                const int64_t c0 = (int64_t)t1 * W1, c1 = std::min<int64_t>(c0 + W1 - 1, NC);
                bool covered = true;
                for (int64_t tx = c0 / OUT; tx <= c1 / OUT && covered; ++tx) covered = tx < ntx2 && plain[tx];
                if (covered) m.h_mask[0][(size_t)fy * ntx1 + t1] = 0;
            }
    }
    for (int d = 1; d < K; ++d)                                   // ring d = ring d-1 dilated by one frame tile
        for (int ty = 0; ty < ntyF; ++ty)
            for (int tx = 0; tx < ntx1; ++tx) {
                if (!m.h_mask[d - 1][(size_t)ty * ntx1 + tx]) continue;
                for (int dy = -1; dy <= 1; ++dy)
                    for (int dx = -1; dx <= 1; ++dx) {
                        const int y = ty + dy, x = tx + dx;
                        if (y >= 0 && y < ntyF && x >= 0 && x < ntx1) m.h_mask[d][(size_t)y * ntx1 + x] = 1;
                    }
            }
    const int w = g.block_warps, nbx = (ntx1 + w - 1) / w;
    for (int d = 0; d < K; ++d)
        for (int ty = 0; ty < ntyF; ++ty)
            for (int bx = 0; bx < nbx; ++bx) {
                bool any = false;
                for (int t = bx * w; t < std::min(ntx1, (bx + 1) * w); ++t) any |= m.h_mask[d][(size_t)ty * ntx1 + t] != 0;
                if (any) m.h_list[d].push_back(make_int2(bx, ty));
            }
}

// Frame mode 3: every tile of the slab is a K-step tile.  Tiles whose K-step cone touches nothing special go to the
// plain kernel (skip map, as above); ALL the others -- the K-row bands at the top / bottom wall, the wall and TF/SF
// columns, tiles near sources, reflectors and probes -- go to stepK_edge, which folds those features into the same
// pipeline as predicates (edge.cuh).  No masks, no pass chain, no scratch generations.  At a slab interface the stored
// halo rows (deep halos, >= K+1 of them) let both kinds of tile reach the slab's first / last owned row.
void plan_edge(const FrameGeom& g, const std::vector<FrameBox>& mods, FrameMaps& m) {
    const int K = g.K, V2 = g.V2;
    const int MARGIN = (K + V2 - 1) / V2, OUT = (32 - 2 * MARGIN) * V2, LD = 32 * V2;
    const int64_t NR = g.rows, NC = g.cols;
    const int64_t lo = g.row_begin, hi = g.row_end + (g.row_end == NR ? 1 : 0);
    const int TR = g.tile_rows;
    const bool top_wall = g.row_begin == 0, bot_wall = g.row_end == NR;
    // plain tiles own rows [loK, hiK): K rows away from a wall of the grid, right up to a slab interface
    int64_t loK = top_wall ? lo + K : lo, hiK = bot_wall ? NR - K : g.row_end;
    if (hiK < loK) { loK = lo; hiK = lo; }
    const int nty = hiK > loK ? (int)((hiK - loK + TR - 1) / TR) : 0;
    const int ntx2 = (int)((NC + 1 + OUT - 1) / OUT);
    int ntx_e = ntx2;                     // a tile whose loaded columns reach past column NC owns everything right of it
    while (ntx_e > 1 && (int64_t)(ntx_e - 2) * OUT - (int64_t)MARGIN * V2 + LD > NC) --ntx_e;
    m.nty = nty; m.ntx2 = ntx2; m.ntx1 = 0; m.frame_tile_rows = TR; m.frame_vec = V2; m.frame_nty = 0;
    m.k_row_lo = loK; m.k_row_hi = hiK;
    m.h_skip.assign((size_t)std::max(nty, 1) * ntx2, 1);
    for (int d = 0; d < 8; ++d) { m.h_mask[d].clear(); m.h_list[d].clear(); }
    std::vector<fdtd::EdgeTile> near, far;
    const int64_t D = std::max<int64_t>(g.halo_top, g.halo_bot);
    // Edge tiles are single warps that march (rows + 2K + 1) iterations of a slow pipeline: they are cut into pieces of
    // at most edge_rows rows, independent of the height of the plain tiles (which like to be tall), so that the few
    // thousand of them spread over the SMs and finish well inside the plain kernel's run.
    const int ER = std::max(g.edge_rows, 1);
    const int64_t first_stored = std::max<int64_t>(g.row_begin - g.halo_top, 0), last_stored = g.row_end + g.halo_bot - 1;
    auto push = [&](int tx, int64_t ta0, int64_t tb0) {
        const int64_t jw = (int64_t)tx * OUT - (int64_t)MARGIN * V2;
        for (int64_t ta = ta0; ta < tb0; ta += ER) {
            const int64_t tb = std::min<int64_t>(ta + ER, tb0);
            const bool is_near = (!top_wall && ta < lo + D) || (!bot_wall && tb > g.row_end - D);
            // what the rows / columns this piece marches through (ta-K-1 .. tb+K, jw .. jw+LD-1) touch: its class
            int need = 0;
            if (!(ta - K - 1 >= std::max<int64_t>(first_stored, 1) && tb + K <= std::min<int64_t>(NR - 2, last_stored))) need |= fdtd::EDGE_ROWS;
            if (jw < 1 || jw + LD - 1 > NC - 2) need |= fdtd::EDGE_COLS;
            for (const FrameBox& q : mods)
                if (q.r0 <= tb + K && q.r1 >= ta - K - 1 && q.c0 <= jw + LD && q.c1 >= jw - 1) need |= q.what;
            (is_near ? near : far).push_back(fdtd::EdgeTile{tx, (int)ta, (int)tb, fdtd::edge_class_for(need)});
        }
    };
    if (loK > lo) for (int tx = 0; tx < ntx_e; ++tx) push(tx, lo, loK);
    for (int ty = 0; ty < nty; ++ty) {
        const int64_t ta = loK + (int64_t)ty * TR, tb = std::min<int64_t>(ta + TR, hiK);
        const bool rows_ok = ta - K >= std::max<int64_t>(g.row_begin - g.halo_top, 0) && ta - (K - 1) >= 1 &&
                             tb + K - 2 <= NR - 2 && tb + K - 1 <= g.row_end + g.halo_bot - 1;
        for (int tx = 0; tx < ntx2; ++tx) {
            const int64_t jw = (int64_t)tx * OUT - (int64_t)MARGIN * V2;
            bool mixed = !rows_ok || jw < 1 || jw + LD - 1 > NC - 2;
            if (!mixed)
                for (const FrameBox& q : mods)
                    if (q.r0 <= tb + K && q.r1 >= ta - K - 1 && q.c0 <= jw + LD && q.c1 >= jw - 1) { mixed = true; break; }
            m.h_skip[(size_t)ty * ntx2 + tx] = mixed;
            if (mixed && tx < ntx_e) push(tx, ta, tb);
        }
    }
    if (hi > hiK && hiK >= loK) for (int tx = 0; tx < ntx_e; ++tx) push(tx, std::max(hiK, lo), hi);
    // tiles of one class are launched together: sort each half by class (stable: DRAM likes the row order kept)
    auto by_class = [](std::vector<fdtd::EdgeTile>& v, std::vector<FrameMaps::EdgeRun>& runs, int base) {
        std::stable_sort(v.begin(), v.end(), [](const fdtd::EdgeTile& a, const fdtd::EdgeTile& b) { return a.pad < b.pad; });
        runs.clear();
        for (size_t k = 0; k < v.size(); ++k) {
            if (runs.empty() || runs.back().flags != v[k].pad) runs.push_back({base + (int)k, 0, v[k].pad});
            runs.back().count++;
        }
    };
    by_class(near, m.edge_runs_near, 0);
    by_class(far, m.edge_runs_far, (int)near.size());
    m.n_edge_near = (int)near.size();
    m.h_edge = near;
    m.h_edge.insert(m.h_edge.end(), far.begin(), far.end());
}

// signature of the frame maps of the group of K calls starting at call n: which pulses are live in any of them, plus the
// geometry knobs
std::vector<uint8_t> frame_key(const fdtd_ctx* c, int64_t n, int K) {
    std::vector<uint8_t> key;
    for (int p = 0; p < c->n_src; ++p) {
        uint8_t live = 0;
        for (int64_t t = n; t < n + K; ++t) if (t < c->table_len && c->active[(size_t)p * c->table_len + t]) live = 1;
        key.push_back(live);
    }
    const bool edge = edge_usable(c, K);
    for (int v : {c->cfg.tile_rows, c->cfg.vec, c->cfg.block_warps, c->n_rect, c->n_probe, K, (int)edge, c->halo_depth}) {
        key.push_back((uint8_t)(v & 0xff)); key.push_back((uint8_t)((v >> 8) & 0xff)); key.push_back((uint8_t)((v >> 16) & 0xff));
    }
    return key;
}

// forced_key: a signature that marks MORE pulses live than this group has (fdtd_run_streamed plans once for all groups of a
// size: a tile near a pulse that is silent in some group is an edge tile there, too -- exact, the kernel reads each call's
// own records).  While c->maps_pinned, existing maps are used as they are.
template <typename T>
int build_frame_maps(fdtd_ctx* c, int64_t n, int K, const std::vector<uint8_t>* forced_key = nullptr) {
    FrameMaps& m = c->fm[K];
    if (c->maps_pinned && !forced_key && m.d_skip && !m.key.empty()) return 0;
    const std::vector<uint8_t> key = forced_key ? *forced_key : frame_key(c, n, K);
    const bool edge = edge_usable(c, K);
    if (key == m.key && m.d_skip) return 0;
    m.band_rows = 0;                                             // the band plan of fdtd_run_streamed follows the maps
    std::vector<FrameBox> mods;
    const int64_t BIG = (int64_t)1 << 40;
    for (int p = 0; p < c->n_src; ++p) {
        if (!key[p]) continue;
        const fdtd_source& q = c->src[p];
        if (q.kind == FDTD_SRC_POINT) mods.push_back({q.row, q.row, q.col, q.col, fdtd::EDGE_SRC});
        else if (q.kind == FDTD_SRC_RIGHT || q.kind == FDTD_SRC_LEFT) mods.push_back({-BIG, BIG, q.col, q.col, fdtd::EDGE_SRC});
        else if (q.kind == FDTD_SRC_UP) mods.push_back({q.row + 1, q.row + 1, -BIG, BIG, fdtd::EDGE_SRC});
        else mods.push_back({q.row, q.row, -BIG, BIG, fdtd::EDGE_SRC});
    }
    for (int r = 0; r < c->n_rect; ++r)
        for (const RectI* b : {&c->rh[r], &c->rex[r], &c->rey[r]})
            if (b->r1 > b->r0 && b->c1 > b->c0) mods.push_back({b->r0, b->r1 - 1, b->c0, b->c1 - 1, fdtd::EDGE_MISC});
    for (int k = 0; k < c->n_probe; ++k) mods.push_back({c->probe_i[k], c->probe_i[k], c->probe_j[k], c->probe_j[k], fdtd::EDGE_MISC});
    const bool deep = c->halo_depth > 1;
    const FrameGeom geom{c->cfg.rows, c->cfg.cols, c->cfg.row_begin, c->cfg.row_end, c->cfg.tile_rows,
                         c->cfg.block_warps, K, kvec<T>(), deep ? c->halo_top : 0, deep ? c->halo_bot : 0, edge,
                         c->edge_rows};
    if (edge) plan_edge(geom, mods, m); else plan_frame(geom, mods, m);
    const int ntyF = m.frame_nty, ntx1 = m.ntx1;
    size_t longest = 1;
    for (int d = 0; d < K; ++d) longest = std::max(longest, m.h_list[d].size());
    // the previous group's launches may still be reading the old maps
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->edge_stream));
    if (m.d_skip_cap < m.h_skip.size()) {
        if (m.d_skip) cudaFree(m.d_skip);
        CK(cudaMalloc(&m.d_skip, m.h_skip.size())); m.d_skip_cap = m.h_skip.size();
    }
    if (m.d_edge_cap < m.h_edge.size()) {
        if (m.d_edge) cudaFree(m.d_edge);
        CK(cudaMalloc(&m.d_edge, m.h_edge.size() * sizeof(fdtd::EdgeTile))); m.d_edge_cap = m.h_edge.size();
    }
    if (edge && !m.h_edge.empty())
        CK(cudaMemcpyAsync(m.d_edge, m.h_edge.data(), m.h_edge.size() * sizeof(fdtd::EdgeTile), cudaMemcpyHostToDevice, c->stream));
    const size_t mask_bytes = (size_t)ntyF * ntx1;
    if (!edge && (m.d_mask_cap < mask_bytes || m.d_list_cap < longest)) {
        for (int d = 0; d < K; ++d) {
            if (m.d_mask[d]) cudaFree(m.d_mask[d]);
            if (m.d_list[d]) cudaFree(m.d_list[d]);
            CK(cudaMalloc(&m.d_mask[d], mask_bytes)); CK(cudaMalloc(&m.d_list[d], longest * sizeof(int2)));
        }
        m.d_mask_cap = mask_bytes; m.d_list_cap = longest;
    }
    CK(cudaMemcpyAsync(m.d_skip, m.h_skip.data(), m.h_skip.size(), cudaMemcpyHostToDevice, c->stream));
    for (int d = 0; d < K && !edge; ++d) {
        CK(cudaMemcpyAsync(m.d_mask[d], m.h_mask[d].data(), mask_bytes, cudaMemcpyHostToDevice, c->stream));
        if (!m.h_list[d].empty())
            CK(cudaMemcpyAsync(m.d_list[d], m.h_list[d].data(), m.h_list[d].size() * sizeof(int2), cudaMemcpyHostToDevice, c->stream));
    }
    CK(cudaStreamSynchronize(c->stream));
    {
        const int V2 = kvec<T>(), OUT = (32 - 2 * ((K + V2 - 1) / V2)) * V2;
        m.row_cells.assign((size_t)std::max(m.nty, 0), 0);
        for (int ty = 0; ty < m.nty; ++ty) {
            const int64_t ta = m.k_row_lo + (int64_t)ty * c->cfg.tile_rows, tb = std::min<int64_t>(ta + c->cfg.tile_rows, m.k_row_hi);
            int plain = 0;
            for (int tx = 0; tx < m.ntx2; ++tx) plain += m.h_skip[(size_t)ty * m.ntx2 + tx] ? 0 : 1;
            m.row_cells[ty] = (int64_t)plain * OUT * (tb - ta);
        }
    }
    m.key = key;
    return 0;
}

// K-step tile rows [ty0, ty0 + ny) of the plan (ny < 0: all of them) on stream st (null: the main stream)
template <typename T, int K>
void launch_stepK(const StepParams<T>& P0, fdtd_ctx* c, int ty0 = 0, int ny = -1, cudaStream_t st = nullptr) {
    constexpr int V2 = kvec<T>();
    const FrameMaps& m = c->fm[K];
    const int w = std::min(c->cfg.block_warps, 4);
    if (ny < 0) ny = m.nty - ty0;
    if (ny <= 0) return;
    if (!st) st = c->stream;
    StepParams<T> P = P0;                                        // the kernel derives a tile's rows from row_lo and blockIdx.y
    P.row_lo = P0.row_lo + ty0 * P0.tile_rows;
    const unsigned char* skip = m.d_skip + (size_t)ty0 * m.ntx2;
    dim3 grid((unsigned)((m.ntx2 + w - 1) / w), (unsigned)ny);
    struct Timed {                                               // event pair around this launch, if asked for (only the
        fdtd_ctx* c; cudaStream_t st; bool on;                   // configured group size: remainder groups are not the headline)
        Timed(fdtd_ctx* c_, cudaStream_t st_, bool want) : c(c_), st(st_),
            on(want && c_->ktime_on && 2 * c_->kt_n + 1 < (int)c_->kt_ev.size()) {
            if (on) cudaEventRecord(c->kt_ev[2 * c->kt_n], st);
        }
        ~Timed() { if (on) { cudaEventRecord(c->kt_ev[2 * c->kt_n + 1], st); c->kt_n++; } }
    } timed(c, st, K == c->temporal);
    if (timed.on)
        for (int ty = ty0; ty < ty0 + ny && ty < (int)m.row_cells.size(); ++ty) c->kt_cell_steps += m.row_cells[ty] * K;
    if constexpr (sizeof(T) == 8) {
        if (c->split) {                                          // split storage: the lean pipeline on hi / lo planes
            fdtd::launch_stepK_split_k<K>(P, w, grid, skip, m.ntx2, st);
            return;
        }
    }
    // 1 = stepK_lean (default; 2 = the same on a uniform mu_coef plane, picked here), 0 = stepK_stream
    const int variant = (c->kstep_variant == 1 && c->mu_uniform) ? (c->skew && K >= 4 ? 3 : 2) : c->kstep_variant;
    fdtd::launch_stepK_k<T, K>(P, variant, w, grid, skip, m.ntx2, st);
}

template <typename T>
void launch_stepK_any(const StepParams<T>& P, fdtd_ctx* c, int K, int ty0 = 0, int ny = -1, cudaStream_t st = nullptr) {
    switch (K) {                                                 // K = 2..8 (fdtd_set_temporal_blocking checks)
        case 2: launch_stepK<T, 2>(P, c, ty0, ny, st); break;
        case 3: launch_stepK<T, 3>(P, c, ty0, ny, st); break;
        case 4: launch_stepK<T, 4>(P, c, ty0, ny, st); break;
        case 5: launch_stepK<T, 5>(P, c, ty0, ny, st); break;
        case 6: launch_stepK<T, 6>(P, c, ty0, ny, st); break;
        case 7: launch_stepK<T, 7>(P, c, ty0, ny, st); break;
        default: launch_stepK<T, 8>(P, c, ty0, ny, st); break;
    }
}

// Steps n .. n+K-1 in one HBM pass: A -> B.  The frame (walls, sources, reflectors, probes, slab interface strips)
// goes A -> S -> S' -> ... -> B with K masked, compact single-step launches: passes 1..K-1 (frame dilated by a ring
// of tiles wide enough for the K-p cells of validity still needed) run on the high-priority edge stream
// CONCURRENTLY with the K-step kernel, the last pass follows both.
// With slabs the generation each pass wrote has its halo rows exchanged on the comm stream right after the pass.
template <typename T>
int group_impl(fdtd_ctx* c, int64_t n, int K) {
    Range nvtx("fdtd:k_group_pass_chain");
    if (int rc = build_frame_maps<T>(c, n, K)) return rc;
    const FrameMaps& m = c->fm[K];
    const int A = c->cur, B = (A + 1) % c->ngen;
    int scratch[2] = {(A + 2) % c->ngen, (A + 3) % c->ngen};
    const int64_t lo = c->cfg.row_begin, hi = c->cfg.row_end + (c->cfg.row_end == c->cfg.rows ? 1 : 0);
    if (c->n_probe) if (int rc = ensure_probe_capacity(c, K)) return rc;
    CK(cudaEventRecord(c->ev_int, c->stream));                  // everything before this group
    CK(cudaStreamWaitEvent(c->edge_stream, c->ev_int, 0));
    int src = A;
    for (int p = 1; p <= K; ++p) {
        const bool last = p == K;
        const int dst = last ? B : scratch[(p - 1) & 1];
        // Pass p must be exact on the frame dilated by K-p CELLS (one cell of validity is lost per step); it runs on
        // the frame dilated by ceil((K-p)/frame_tile_rows) TILES.  Values it computes further out from stale inputs
        // are never read by a value that is kept.
        const int ring = (K - p + m.frame_tile_rows - 1) / m.frame_tile_rows;
        // The last pass is independent of the K-step kernel too (it writes frame tiles of B from a scratch generation),
        // but queuing it on the edge stream next to the kernel measured SLOWER (433.8 vs 450.0 Gcell/s on the C5 slab,
        // profiles/r1_overlap_last_pass.json: its latency-bound CTAs take SM slots from the kernel for longer than the
        // tail they save), so it follows the kernel on the main stream.
        const bool on_edge = !last;
        cudaStream_t st = on_edge ? c->edge_stream : c->stream;
        if (c->comm) CK(cudaStreamWaitEvent(st, c->ev_comm, 0));                  // halos of `src` have landed
        if (last) {
            StepParams<T> PK;                                   // the K-step kernel: everything but the frame
            if (int rc = fill_params<T>(c, PK, n, A, B)) return rc;
            PK.n_src = 0; PK.n_rect = 0; PK.n_probe = 0; PK.row_lo = (int)m.k_row_lo; PK.row_hi = (int)m.k_row_hi;
            if (m.nty > 0) launch_stepK_any<T>(PK, c, K);
            if (m.nty > 0) c->stats.kernel_launches++;
            CK(cudaGetLastError());
            if (!on_edge) CK(cudaStreamWaitEvent(c->stream, c->ev_bnd, 0));       // passes 1..K-1 done
        }
        StepParams<T> P;
        if (int rc = fill_params<T>(c, P, n + p - 1, src, dst)) return rc;
        if (c->n_probe) P.probe_out = c->probe_dev + (size_t)(c->probe_count + p - 1) * c->n_probe * 3;
        P.tile_mask = m.d_mask[ring]; P.mask_stride = m.ntx1; P.tile_list = m.d_list[ring];
        P.tile_rows = m.frame_tile_rows;
        if (int rc = launch_rows<T>(c, P, lo, hi, st, (int)m.h_list[ring].size(), m.frame_vec)) return rc;
        if (on_edge) CK(cudaEventRecord(c->ev_bnd, c->edge_stream));
        if (c->comm) {
            cudaEvent_t done = on_edge ? c->ev_bnd : c->ev_p3;
            if (!on_edge) CK(cudaEventRecord(c->ev_p3, c->stream));
            CK(cudaStreamWaitEvent(c->comm_stream, done, 0));
            if (int rc = exchange(c, dst, c->comm_stream)) return rc;
            CK(cudaEventRecord(c->ev_comm, c->comm_stream));
        }
        src = dst;
    }
    CK(cudaEventRecord(c->ev_int, c->stream));                  // "all owned rows of generation B written"
    c->cur = B;
    if (c->n_probe) c->probe_count += K;
    c->stats.steps += K;
    c->stats.cell_updates += (int64_t)K * c->owned * c->cfg.cols;
    return 0;
}

// ---- resident mode: many steps per cooperative launch (resident.cuh) ------------------------------------
bool resident_eligible(const fdtd_ctx* c) {
    if (!c->resident || c->comm || c->cfg.kernel != FDTD_KERNEL_STREAM || c->split) return false;
    if (c->cfg.row_begin != 0 || c->cfg.row_end != c->cfg.rows) return false;            // whole grid on this GPU
    if ((c->cfg.rows + 1) * (c->cfg.cols + 1) > c->resident_max_cells) return false;
    if (c->n_src > FDTD_MAX_SRC) return false;
    for (int p = 0; p < c->n_src; ++p)
        if (c->src[p].kind == FDTD_SRC_RIGHT && c->src[p].col == 0) return false;        // h[:, col-1] wraps to the far column
    return true;
}

void resident_free_tables(fdtd_ctx* c) {
    for (void** p : {&c->rs_count, &c->rs_last, &c->rs_src}) { if (*p) cudaFree(*p); *p = nullptr; }
    c->rs_valid = false; c->rs_stride = 0;
}

// Per-call records of the ACTIVE pulses in list order (what fill_params builds for one call), for every call of the
// tables, uploaded once per fdtd_set_sources.
template <typename T>
int resident_tables(fdtd_ctx* c) {
    if (c->rs_valid) return 0;
    CK(cudaStreamSynchronize(c->stream));                        // an earlier launch may still read the old records
    resident_free_tables(c);
    if (c->n_src > 0) {
        const int64_t L = c->table_len;
        const int stride = c->n_src;
        std::vector<int> count((size_t)L, 0);
        std::vector<T> last((size_t)L, T(0));
        std::vector<fdtd::SrcStep<T>> src((size_t)L * stride);
        std::memset(src.data(), 0, src.size() * sizeof(fdtd::SrcStep<T>));
        for (int64_t n = 0; n < L; ++n) {
            int k = 0;
            for (int p = 0; p < c->n_src; ++p) {
                if (!c->active[(size_t)p * L + n]) continue;
                fdtd::SrcStep<T>& q = src[(size_t)n * stride + k++];
                q.kind = c->src[p].kind; q.row = (int)c->src[p].row; q.col = (int)c->src[p].col; q.pad = 0;
                q.mag = (T)c->mag[(size_t)p * L + n];
                q.magz = (T)c->magz[(size_t)p * L + n];
            }
            count[(size_t)n] = k;
            last[(size_t)n] = (T)c->last[(size_t)n];
        }
        CK(cudaMalloc(&c->rs_count, count.size() * sizeof(int)));
        CK(cudaMalloc(&c->rs_last, last.size() * sizeof(T)));
        CK(cudaMalloc(&c->rs_src, src.size() * sizeof(fdtd::SrcStep<T>)));
        CK(cudaMemcpy(c->rs_count, count.data(), count.size() * sizeof(int), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->rs_last, last.data(), last.size() * sizeof(T), cudaMemcpyHostToDevice));
        CK(cudaMemcpy(c->rs_src, src.data(), src.size() * sizeof(fdtd::SrcStep<T>), cudaMemcpyHostToDevice));
        c->rs_stride = stride;
    }
    c->rs_valid = true;
    return 0;
}

// calls n .. n+g-1 in ONE cooperative launch
template <typename T>
int resident_run(fdtd_ctx* c, int64_t n, int64_t g) {
    Range nvtx("fdtd:resident_run");
    using G = fdtd::ResGeom;
    if (int rc = resident_tables<T>(c)) return rc;
    const int A = c->cur, B = next_gen(c, A);
    StepParams<T> P;
    if (int rc = fill_params<T>(c, P, n, A, B)) return rc;
    P.n_src = 0; P.last_mag = T(0);                              // the kernel loads each call's record itself
    P.row_lo = 0; P.row_hi = (int)c->cfg.rows + 1;
    if (c->n_probe) P.probe_out = c->probe_dev + (size_t)c->probe_count * c->n_probe * 3;
    if (!c->res_blocks_max) {
        int coop = 0, sms = 0, per_sm = 0;
        CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, c->cfg.device));
        CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, c->cfg.device));
        CK(fdtd::resident_blocks_per_sm<T>(&per_sm));
        if (!coop || per_sm < 1) return fail(c, FDTD_ERR_CUDA, "cooperative launch unavailable for step_resident");
        c->res_blocks_max = per_sm * sms;
    }
    const int tiles = fdtd::resident_tiles_x<G>(P.NC) * fdtd::resident_tiles_y<G>(P.NR);
    const int blocks = std::min(tiles, c->res_blocks_max);       // all CTAs co-resident; they stride over the tiles
    T* h1 = (T*)c->gen[B][0]; T* ex1 = (T*)c->gen[B][1]; T* ey1 = (T*)c->gen[B][2];
    fdtd::ResSources<T> R{(const int*)c->rs_count, (const T*)c->rs_last, (const fdtd::SrcStep<T>*)c->rs_src, c->rs_stride};
    CK(fdtd::launch_step_resident<T>(P, h1, ex1, ey1, R, n, (int)g, blocks, c->stream));
    c->stats.kernel_launches++;
    c->cur = (g & 1) ? B : A;
    if (c->n_probe) c->probe_count += g;
    c->stats.steps += g;
    c->stats.cell_updates += g * c->owned * c->cfg.cols;
    return 0;
}

int resident_any(fdtd_ctx* c, int64_t n, int64_t g) {
    return c->cfg.dtype == FDTD_F64 ? resident_run<double>(c, n, g) : resident_run<float>(c, n, g);
}

// ---- frame mode 3: plain K-step tiles + stepK_edge tiles, nothing else ------------------------------------------------
inline cudaError_t launch_edge_split_any(int K, const StepParams<double>& P, const fdtd::EdgeTile* tiles, int n,
                                         const fdtd::ResSources<double>& R, long long first, cudaStream_t st) {
    switch (K) {
        case 1: return fdtd::launch_edge_split_k<1>(P, tiles, n, R, first, st);
        case 2: return fdtd::launch_edge_split_k<2>(P, tiles, n, R, first, st);
        case 3: return fdtd::launch_edge_split_k<3>(P, tiles, n, R, first, st);
        case 4: return fdtd::launch_edge_split_k<4>(P, tiles, n, R, first, st);
        case 5: return fdtd::launch_edge_split_k<5>(P, tiles, n, R, first, st);
        case 6: return fdtd::launch_edge_split_k<6>(P, tiles, n, R, first, st);
        case 7: return fdtd::launch_edge_split_k<7>(P, tiles, n, R, first, st);
        case 8: return fdtd::launch_edge_split_k<8>(P, tiles, n, R, first, st);
        default: return cudaErrorInvalidValue;
    }
}

template <typename T>
cudaError_t launch_edge_any(int K, const StepParams<T>& P, int flags, const fdtd::EdgeTile* tiles, int n,
                            const fdtd::ResSources<T>& R, long long first, cudaStream_t st) {
    if constexpr (sizeof(T) == 8) {
        // split storage (P.split_lo_off != 0) has ONE edge instantiation, the one with every feature
        if (P.split_lo_off) return launch_edge_split_any(K, P, tiles, n, R, first, st);
    }
    switch (K) {
        case 2: return fdtd::launch_edge_k<T, 2>(P, flags, tiles, n, R, first, st);
        case 3: return fdtd::launch_edge_k<T, 3>(P, flags, tiles, n, R, first, st);
        case 4: return fdtd::launch_edge_k<T, 4>(P, flags, tiles, n, R, first, st);
        case 5: return fdtd::launch_edge_k<T, 5>(P, flags, tiles, n, R, first, st);
        case 6: return fdtd::launch_edge_k<T, 6>(P, flags, tiles, n, R, first, st);
        case 7: return fdtd::launch_edge_k<T, 7>(P, flags, tiles, n, R, first, st);
        default: return fdtd::launch_edge_k<T, 8>(P, flags, tiles, n, R, first, st);
    }
}

// the edge tiles of the plan next to a slab interface (near) or all the others: one launch per tile class in that half
template <typename T>
int launch_edge(fdtd_ctx* c, int64_t n, int K, int A, int B, bool near_half, cudaStream_t st) {
    const FrameMaps& m = c->fm[K];
    const std::vector<FrameMaps::EdgeRun>& runs = near_half ? m.edge_runs_near : m.edge_runs_far;
    if (runs.empty()) return 0;
    if (int rc = resident_tables<T>(c)) return rc;
    StepParams<T> PE;
    if (int rc = fill_params<T>(c, PE, n, A, B)) return rc;
    PE.n_src = 0; PE.last_mag = T(0);                            // the kernel reads the K calls' records itself
    PE.row_lo = (int)c->cfg.row_begin;
    PE.row_hi = (int)(c->cfg.row_end + (c->cfg.row_end == c->cfg.rows ? 1 : 0));
    if (c->n_probe) PE.probe_out = c->probe_dev + (size_t)c->probe_count * c->n_probe * 3;
    fdtd::ResSources<T> R{(const int*)c->rs_count, (const T*)c->rs_last, (const fdtd::SrcStep<T>*)c->rs_src, c->rs_stride};
    for (const FrameMaps::EdgeRun& run : runs) {                 // costliest class first: it is the one with the longest tiles
        CK(launch_edge_any<T>(K, PE, c->edge_class_force ? fdtd::EDGE_ALL : run.flags, m.d_edge + run.first, run.count, R, n, st));
        c->stats.kernel_launches++;
    }
    return 0;
}

// Steps n .. n+K-1: A -> B in two launches that read only A and write disjoint tiles of B.  The edge tiles (long-running
// single warps) start first, on the high-priority stream, next to the plain kernel.  On a y-slab with a communicator the
// tiles next to an interface (plain tile rows and edge tiles within D rows of it) run on that stream, too, once the
// previous exchange has landed; the D-row blocks of B leave as soon as they are done, while the interior still runs.
template <typename T>
int group_edge(fdtd_ctx* c, int64_t n, int K) {
    Range nvtx("fdtd:k_group");
    if (int rc = build_frame_maps<T>(c, n, K)) return rc;
    const FrameMaps& m = c->fm[K];
    const int A = c->cur, B = (A + 1) % c->ngen;
    if (c->n_probe) if (int rc = ensure_probe_capacity(c, K)) return rc;
    StepParams<T> PK;
    if (m.nty > 0) {
        if (int rc = fill_params<T>(c, PK, n, A, B)) return rc;
        PK.n_src = 0; PK.n_rect = 0; PK.n_probe = 0; PK.row_lo = (int)m.k_row_lo; PK.row_hi = (int)m.k_row_hi;
    }
    const int n_edge = (int)m.h_edge.size();
    CK(cudaEventRecord(c->ev_int, c->stream));                  // everything before this group
    CK(cudaStreamWaitEvent(c->edge_stream, c->ev_int, 0));
    if (!c->comm) {
        if (int rc = launch_edge<T>(c, n, K, A, B, true, c->edge_stream)) return rc;
        if (int rc = launch_edge<T>(c, n, K, A, B, false, c->edge_stream)) return rc;
        if (m.nty > 0) { launch_stepK_any<T>(PK, c, K); c->stats.kernel_launches++; CK(cudaGetLastError()); }
    } else {
        const int TR = c->cfg.tile_rows, D = c->halo_depth;
        const int e_top = c->cfg.row_begin > 0 ? std::min(m.nty, (D + TR - 1) / TR) : 0;
        const int e_bot = c->cfg.row_end < c->cfg.rows ? std::min(m.nty - e_top, (D + TR - 1) / TR + 1) : 0;
        const int n_int = m.nty - e_top - e_bot;
        // tiles away from the interfaces read owned rows only: they start at once
        if (int rc = launch_edge<T>(c, n, K, A, B, false, c->edge_stream)) return rc;
        if (n_int > 0) { launch_stepK_any<T>(PK, c, K, e_top, n_int, c->stream); c->stats.kernel_launches++; }
        CK(cudaStreamWaitEvent(c->edge_stream, c->ev_comm, 0)); // the neighbours' rows of A have landed
        if (int rc = launch_edge<T>(c, n, K, A, B, true, c->edge_stream)) return rc;
        if (e_top > 0) { launch_stepK_any<T>(PK, c, K, 0, e_top, c->edge_stream); c->stats.kernel_launches++; }
        if (e_bot > 0) { launch_stepK_any<T>(PK, c, K, m.nty - e_bot, e_bot, c->edge_stream); c->stats.kernel_launches++; }
        CK(cudaGetLastError());
    }
    CK(cudaEventRecord(c->ev_bnd, c->edge_stream));
    if (c->comm) {
        CK(cudaStreamWaitEvent(c->comm_stream, c->ev_bnd, 0));
        if (int rc = exchange(c, B, c->comm_stream)) return rc; // ONE exchange per group
        CK(cudaEventRecord(c->ev_comm, c->comm_stream));
    }
    CK(cudaStreamWaitEvent(c->stream, c->ev_bnd, 0));           // main-stream order = "all of B written"
    CK(cudaEventRecord(c->ev_int, c->stream));
    c->cur = B;
    if (c->n_probe) c->probe_count += K;
    c->stats.steps += K;
    c->stats.cell_updates += (int64_t)K * c->owned * c->cfg.cols;
    return 0;
}

// ---- split field storage (dtype FDTD_F32X) ----------------------------------------------------------------------------------
// Fields: fp32 hi plane + bf16 lo plane per buffer (6 bytes per value, 32 significant bits); coefficients, pulse tables and
// arithmetic: float64, in the reference's operation order.  One kernel family runs this mode: stepK_edge with every
// feature (edge.cuh, SPLIT = 1), for every tile of the grid, K = 1 .. 8 calls per launch with the bits of single calls.
float* split_hi(fdtd_ctx* c, void* base, int64_t global_row) { return (float*)base + (global_row - c->row_off) * c->pitch; }
unsigned short* split_lo(fdtd_ctx* c, void* base, int64_t global_row) {
    return (unsigned short*)((char*)base + (size_t)c->lrows * c->pitch * 4) + (global_row - c->row_off) * c->pitch;
}

int split_h2d_field(fdtd_ctx* c, void* base, const double* host, int64_t host_ld, int64_t g0, int64_t n, int64_t ncols) {
    if (n <= 0 || !host) return 0;
    const double* src = host + (g0 - c->host_row0) * host_ld;
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(32 << 20) / (ncols * 8));
    if (int rc = ensure_stage(c, (size_t)std::min(chunk, n) * ncols * 8)) return rc;
    for (int64_t a = 0; a < n; a += chunk) {
        const int64_t m = std::min(chunk, n - a);
        CK(cudaMemcpy2DAsync(c->stage, ncols * 8, src + a * host_ld, host_ld * 8, ncols * 8, m, cudaMemcpyHostToDevice, c->stream));
        dim3 grid((unsigned)((ncols + 255) / 256), (unsigned)std::min<int64_t>(m, 65535));
        CK(fdtd::launch_split_from_f64(grid, c->stream, split_hi(c, base, g0 + a), split_lo(c, base, g0 + a), c->pitch,
                                       (const double*)c->stage, ncols, (int)m, (int)ncols));
        c->stats.kernel_launches++;
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

int split_d2h_field(fdtd_ctx* c, void* base, double* host, int64_t host_ld, int64_t g0, int64_t n, int64_t ncols) {
    if (n <= 0 || !host) return 0;
    double* dst = host + (g0 - c->host_row0) * host_ld;
    const int64_t chunk = std::max<int64_t>(1, (int64_t)(32 << 20) / (ncols * 8));
    if (int rc = ensure_stage(c, (size_t)std::min(chunk, n) * ncols * 8)) return rc;
    fdtd::PatchList none; none.n = 0;
    for (int64_t a = 0; a < n; a += chunk) {
        const int64_t m = std::min(chunk, n - a);
        dim3 block(64, 4), grid((unsigned)((ncols + 63) / 64), (unsigned)std::min<int64_t>((m + 3) / 4, 65535));
        CK(fdtd::launch_split_to_f64(grid, block, c->stream, (double*)c->stage, ncols, split_hi(c, base, g0 + a),
                                     split_lo(c, base, g0 + a), c->pitch, (int)m, (int)ncols, 1, 1, 0, none));
        c->stats.kernel_launches++;
        CK(cudaMemcpy2DAsync(dst + a * host_ld, host_ld * 8, c->stage, ncols * 8, ncols * 8, m, cudaMemcpyDeviceToHost, c->stream));
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// the tiles of a group of K calls: the whole index space (rows 0 .. NR, columns 0 .. NC), every tile with every feature
int split_plan(fdtd_ctx* c, int K) {
    fdtd_ctx::SplitTiles& t = c->split_tiles[K];
    if (t.dev) return 0;
    const int V2 = 2, MARGIN = (K + V2 - 1) / V2, OUT = (32 - 2 * MARGIN) * V2, LD = 32 * V2;
    const int64_t NR = c->cfg.rows, NC = c->cfg.cols;
    int ntx = (int)((NC + 1 + OUT - 1) / OUT);
    while (ntx > 1 && (int64_t)(ntx - 2) * OUT - (int64_t)MARGIN * V2 + LD > NC) --ntx;    // that tile owns everything right of it
    // tile height: a tile marches rows + 2 K + 1 iterations of a latency-bound pipeline, so short tiles waste work and
    // tall ones leave a small grid with too few warps; aim at about four warps per resident-warp slot of the GPU
    int64_t er = (NR + 1) * ntx / (4 * 1184);
    er = std::max<int64_t>(8, std::min<int64_t>(64, er));
    if (const char* e = std::getenv("FDTD_SPLIT_ROWS")) { const int v = std::atoi(e); if (v >= 1 && v <= 4096) er = v; }
    std::vector<fdtd::EdgeTile> tiles;
    for (int64_t ta = 0; ta < NR + 1; ta += er)
        for (int tx = 0; tx < ntx; ++tx)
            tiles.push_back(fdtd::EdgeTile{tx, (int)ta, (int)std::min<int64_t>(ta + er, NR + 1), fdtd::EDGE_ALL});
    CK(cudaMalloc(&t.dev, tiles.size() * sizeof(fdtd::EdgeTile)));
    CK(cudaMemcpy(t.dev, tiles.data(), tiles.size() * sizeof(fdtd::EdgeTile), cudaMemcpyHostToDevice));
    t.n = (int)tiles.size();
    return 0;
}

// calls n .. n+K-1 (K = 1 .. 8): generation A -> B in one launch
int split_group(fdtd_ctx* c, int64_t n, int K) {
    Range nvtx("fdtd:split_group");
    for (int p = 0; p < c->n_src; ++p)
        if (c->src[p].kind == FDTD_SRC_RIGHT && c->src[p].col == 0)
            return fail(c, FDTD_ERR_ARG, "split storage (FDTD_F32X) does not take a TF/SF \"right\" line at column 0");
    if (c->n_src > FDTD_MAX_SRC) return fail(c, FDTD_ERR_ARG, "split storage (FDTD_F32X) takes at most 16 pulses");
    if (int rc = split_plan(c, K)) return rc;
    if (int rc = resident_tables<double>(c)) return rc;
    if (c->n_probe) if (int rc = ensure_probe_capacity(c, K)) return rc;
    const int A = c->cur, B = (A + 1) % 2;
    StepParams<double> P;
    if (int rc = fill_params<double>(c, P, n, A, B)) return rc;
    P.n_src = 0; P.last_mag = 0.0;                               // the kernel reads each call's records itself
    P.row_lo = 0; P.row_hi = (int)c->cfg.rows + 1;
    if (c->n_probe) P.probe_out = c->probe_dev + (size_t)c->probe_count * c->n_probe * 3;
    fdtd::ResSources<double> R{(const int*)c->rs_count, (const double*)c->rs_last, (const fdtd::SrcStep<double>*)c->rs_src, c->rs_stride};
    const fdtd_ctx::SplitTiles& t = c->split_tiles[K];
    if (K < 1 || K > 8) return fail(c, FDTD_ERR_ARG, "split group of 1..8 calls");
    cudaError_t e = launch_edge_split_any(K, P, t.dev, t.n, R, n, c->stream);
    CK(e);
    c->stats.kernel_launches++;
    c->cur = B;
    if (c->n_probe) c->probe_count += K;
    c->stats.steps += K;
    c->stats.cell_updates += (int64_t)K * c->owned * c->cfg.cols;
    return 0;
}

int step_any(fdtd_ctx* c, int64_t step) {
    if (c->split) return split_group(c, step, 1);
    return c->cfg.dtype == FDTD_F64 ? step_impl<double>(c, step) : step_impl<float>(c, step);
}

int group_any(fdtd_ctx* c, int64_t n, int K) {
    // split storage: plain tiles through stepK_lean_split, the others through stepK_edge<.., split> (group_edge); a scene
    // stepK_edge cannot take ends in split_group, which says so
    if (c->split) return (edge_usable(c, K) && !c->split_all_edge) ? group_edge<double>(c, n, K) : split_group(c, n, K);
    const bool whole = c->cfg.row_begin == 0 && c->cfg.row_end == c->cfg.rows;
    if (edge_usable(c, K)) return c->cfg.dtype == FDTD_F64 ? group_edge<double>(c, n, K) : group_edge<float>(c, n, K);
    // The fallback, K masked single-step passes through the scratch generations (frame mode 0; a scene stepK_edge cannot
    // take), exchanges one row per pass on a y-slab: it exists for the one-row halo layout only and needs the scratch.
    const int kmax = c->ngen >= 4 ? 8 : (c->ngen == 3 ? 2 : 1);
    if ((c->halo_depth > 1 && !whole) || K > kmax) {
        for (int k = 0; k < K; ++k)
            if (int rc = step_any(c, n + k)) return rc;
        return 0;
    }
    return c->cfg.dtype == FDTD_F64 ? group_impl<double>(c, n, K) : group_impl<float>(c, n, K);
}

// pinned ring slot -> the caller's (pageable) snapshot array, on the driver's callback thread.  A C5-sized snapshot is
// 4.3 GB whose destination pages are touched for the first time here: one memcpy thread moves about 4 GB/s, far below the
// 50 GB/s the DMA before it delivers, so big copies are cut into 2 MiB-aligned chunks for up to 8 threads
// (FDTD_SNAPSHOT_THREADS).  The reference scripts' snapshots (0.15 - 0.8 MB) stay a plain memcpy.
void snapshot_copy(void* dst, const void* src, size_t bytes) {
    static const unsigned want = [] {
        unsigned n = std::min(8u, std::max(1u, std::thread::hardware_concurrency()));
        if (const char* e = std::getenv("FDTD_SNAPSHOT_THREADS")) { const int v = std::atoi(e); if (v >= 1 && v <= 64) n = (unsigned)v; }
        return n;
    }();
    const size_t min_chunk = (size_t)32 << 20;
    const unsigned n = (unsigned)std::min<size_t>(want, bytes / min_chunk);
    if (n <= 1) { std::memcpy(dst, src, bytes); return; }
    const size_t chunk = ((bytes / n) + ((size_t)2 << 20) - 1) & ~(((size_t)2 << 20) - 1);
    std::vector<std::thread> pool;
    for (unsigned t = 1; t < n; ++t) {
        const size_t lo = std::min(bytes, t * chunk), hi = std::min(bytes, (t + 1) * chunk);
        if (hi > lo) pool.emplace_back([=] { std::memcpy((char*)dst + lo, (const char*)src + lo, hi - lo); });
    }
    std::memcpy(dst, src, std::min(bytes, chunk));
    for (std::thread& th : pool) th.join();
}

template <typename T>
int snapshot_impl(fdtd_ctx* c, int slot, double* host_dst, int64_t next_step, int64_t total_calls) {
    Range nvtx("fdtd:snapshot");
    fdtd::PatchList pl; pl.n = 0;
    if (next_step < total_calls && next_step < c->table_len) {
        for (int p = 0; p < c->n_src; ++p) {
            if (c->src[p].kind != FDTD_SRC_POINT || !c->active[(size_t)p * c->table_len + next_step]) continue;
            if (pl.n >= FDTD_MAX_SRC) return fail(c, FDTD_ERR_ARG, "too many point sources for a snapshot");
            pl.p[pl.n].row = (int)c->src[p].row; pl.p[pl.n].col = (int)c->src[p].col;
            // the stored array has dtype T: the reference writes into the array itself
            pl.p[pl.n].value = (double)(T)c->mag[(size_t)p * c->table_len + next_step];
            pl.n++;
        }
    }
    const int64_t rows = c->owned, cols = c->cfg.cols;
    CK(cudaStreamWaitEvent(c->stream, c->ev_free[slot], 0));
    dim3 block(64, 4), grid((unsigned)((cols + 63) / 64), (unsigned)std::min<int64_t>((rows + 3) / 4, 65535));
    if (c->split) {
        CK(fdtd::launch_split_to_f64(grid, block, c->stream, c->snap_dev[slot], cols, split_hi(c, c->gen[c->cur][0], c->cfg.row_begin),
                                     split_lo(c, c->gen[c->cur][0], c->cfg.row_begin), c->pitch, (int)rows, (int)cols, 1, 1,
                                     (int)c->cfg.row_begin, pl));
    } else
    CK(fdtd::launch_plane_to_f64<T>(grid, block, c->stream, c->snap_dev[slot], cols,
        plane<T>(c->gen[c->cur][0], c, c->cfg.row_begin), c->pitch, (int)rows, (int)cols, (int)c->cfg.row_begin, pl));
    c->stats.kernel_launches++;
    CK(cudaEventRecord(c->ev_packed[slot], c->stream));
    CK(cudaStreamWaitEvent(c->copy_stream, c->ev_packed[slot], 0));
    // device slot -> pinned host slot (async DMA) -> the caller's array (plain memcpy run by the driver's callback
    // thread, in stream order).  The caller's array can stay pageable: pinning a whole snapshot stack costs ~1 ms per MB
    // (0.3 s for the tutorial's 331 MB, seconds for doubleslit's 5.2 GB), more than the run itself.
    CK(cudaMemcpyAsync(c->snap_pin[slot], c->snap_dev[slot], (size_t)rows * cols * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    // one heap job per snapshot: the host enqueues far ahead of the callbacks, so a per-slot job would be overwritten
    auto* job = new fdtd_ctx::SnapCopy{host_dst, c->snap_pin[slot], (size_t)rows * cols * 8};
    cudaError_t he = cudaLaunchHostFunc(c->copy_stream, [](void* p) {
        auto* j = static_cast<fdtd_ctx::SnapCopy*>(p);
        snapshot_copy(j->dst, j->src, j->bytes);
        delete j;
    }, job);
    if (he != cudaSuccess) { delete job; return fail(c, FDTD_ERR_CUDA, std::string("cudaLaunchHostFunc: ") + cudaGetErrorString(he)); }
    CK(cudaEventRecord(c->ev_free[slot], c->copy_stream));
    return 0;
}

}  // namespace

// ------------------------------------------------------------------------------------------------
extern "C" {

int64_t fdtd_default_pitch(int64_t cols, int32_t /*dtype*/) { return ((cols + 1 + 31) / 32) * 32; }

int64_t fdtd_local_rows(const fdtd_config* cfg) {
    if (!cfg) return 0;
    return (cfg->row_end - cfg->row_begin) + (cfg->row_begin > 0 ? 1 : 0) + 1;
}

int64_t fdtd_local_rows_halo(const fdtd_config* cfg, int32_t halo_depth) {
    if (!cfg) return 0;
    if (halo_depth <= 1) return fdtd_local_rows(cfg);
    return (cfg->row_end - cfg->row_begin) + (cfg->row_begin > 0 ? halo_depth : 0) +
           (cfg->row_end < cfg->rows ? halo_depth : 0) + 1;
}

const char* fdtd_last_error(fdtd_handle h) { return h ? h->err.c_str() : g_create_error.c_str(); }

static void auto_geometry(fdtd_ctx* c);
static int quiesce(fdtd_ctx* c);

int fdtd_create(const fdtd_config* cfg, fdtd_handle* out) {
    fdtd_ctx* c = nullptr;
    if (!cfg || !out) return fail(c, FDTD_ERR_ARG, "null argument");
    *out = nullptr;
    if (cfg->rows < 3 || cfg->cols < 3) return fail(c, FDTD_ERR_ARG, "grid must be at least 3x3");
    if (cfg->rows > 0x7ffffff0LL || cfg->cols > 0x7ffffff0LL) return fail(c, FDTD_ERR_ARG, "grid too large");
    if (cfg->row_begin < 0 || cfg->row_end > cfg->rows || cfg->row_end - cfg->row_begin < 1)
        return fail(c, FDTD_ERR_ARG, "bad slab row range");
    const bool whole = cfg->row_begin == 0 && cfg->row_end == cfg->rows;
    if (!whole && cfg->row_end - cfg->row_begin < 2) return fail(c, FDTD_ERR_ARG, "a slab must own at least 2 rows");
    if (cfg->dtype != FDTD_F64 && cfg->dtype != FDTD_F32 && cfg->dtype != FDTD_F32X) return fail(c, FDTD_ERR_ARG, "bad dtype");
    if (cfg->dtype == FDTD_F32X && !whole)
        return fail(c, FDTD_ERR_ARG, "split storage (FDTD_F32X) runs on one GPU: y-slabs are not available in this mode");
    if (cfg->kernel != FDTD_KERNEL_EXACT && cfg->kernel != FDTD_KERNEL_STREAM) return fail(c, FDTD_ERR_ARG, "bad kernel");
    int ndev = 0;
    cudaError_t e = cudaGetDeviceCount(&ndev);
    if (e != cudaSuccess || ndev == 0)
        return fail(c, FDTD_ERR_CUDA, std::string("no CUDA device: ") + cudaGetErrorString(e) +
                                          " (libfdtd2d has no CPU path)");
    if (cfg->device < 0 || cfg->device >= ndev) return fail(c, FDTD_ERR_ARG, "bad device ordinal");
    c = new fdtd_ctx();
    c->cfg = *cfg;
    if (const char* eg = std::getenv("FDTD_EDGE_GENERIC")) c->edge_class_force = std::atoi(eg) != 0;
    if (const char* er = std::getenv("FDTD_EDGE_ROWS")) { const int v = std::atoi(er); if (v >= 1 && v <= 4096) c->edge_rows = v; }
    c->esize = cfg->dtype == FDTD_F64 ? 8 : (cfg->dtype == FDTD_F32X ? 6 : 4);     // bytes per FIELD value (F32X: 4 hi + 2 lo)
    c->split = cfg->dtype == FDTD_F32X;
    if (const char* e = std::getenv("FDTD_SPLIT_ALL_EDGE")) c->split_all_edge = std::atoi(e) != 0;
    if (const char* e = std::getenv("FDTD_UNIFORM_MU")) c->mu_uniform_enable = std::atoi(e) != 0;
    if (const char* e = std::getenv("FDTD_SKEW")) c->skew = std::atoi(e) != 0;
    c->pitch = cfg->pitch ? cfg->pitch : fdtd_default_pitch(cfg->cols, cfg->dtype);
    if (c->pitch < cfg->cols + 1 || c->pitch % 32) { delete c; c = nullptr; return fail(c, FDTD_ERR_ARG, "pitch must be >= cols+1 and a multiple of 32"); }
    c->cfg.pitch = c->pitch;
    c->halo_top = cfg->row_begin > 0 ? 1 : 0;
    c->row_off = cfg->row_begin - c->halo_top;
    c->owned = cfg->row_end - cfg->row_begin;
    c->lrows = fdtd_local_rows(cfg);
    {
        // Launch geometry defaults (tools/sweep.py on the 8192 x 65536 slab): 256-bit lanes, 4 warps per CTA and
        // SHORT tiles -- 6 rows keep all resident CTAs inside a narrow band of rows, which is what DRAM likes.
        // Small grids (the script configs, N = 139..442) get narrow single-warp tiles so that the launch
        // still spreads over the 148 SMs.
        const int va = cfg->dtype != FDTD_F32 ? 2 : 4;
        const int64_t wide_tiles = ((cfg->cols + 1 + 128 * va - 1) / (128 * va)) * ((c->owned + 5) / 6);
        const bool small = wide_tiles < 2 * 148;
        if (c->cfg.vec <= 0) c->cfg.vec = small ? va : 4;     // 4 columns per lane: 256-bit (f64) / 128-bit (f32) accesses
        c->auto_warps = c->cfg.block_warps <= 0 && !small;
        c->auto_tile_rows = c->cfg.tile_rows <= 0 && !small;
        if (c->cfg.block_warps <= 0) c->cfg.block_warps = small ? 1 : 4;
        if (c->cfg.tile_rows <= 0) {
            c->cfg.tile_rows = 6;
            if (small) {
                const int64_t col_tiles = (cfg->cols + 1 + 32 * va - 1) / (32 * va);
                const int64_t want_row_tiles = std::max<int64_t>(1, 592 / col_tiles);
                c->cfg.tile_rows = (int)std::min<int64_t>(16, std::max<int64_t>(4, c->owned / want_row_tiles));
            }
        }
    }
    if (c->split) {
        // split storage always runs groups of up to 8 calls per launch (a group has the bits of single calls and needs no
        // extra buffers), so its launch geometry is the K = 8 one from the start; fdtd_set_temporal_blocking may lower K
        c->temporal = 8;
        auto_geometry(c);
    }
    auto bail = [&](cudaError_t ee, const char* what) {
        std::string m = std::string(what) + ": " + cudaGetErrorString(ee);
        delete c; c = nullptr; return fail(nullptr, FDTD_ERR_CUDA, m);
    };
    if ((e = cudaSetDevice(cfg->device)) != cudaSuccess) return bail(e, "cudaSetDevice");
    if ((e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithFlags(&c->copy_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    int prio_lo = 0, prio_hi = 0;
    cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
    if ((e = cudaStreamCreateWithPriority(&c->comm_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    if ((e = cudaStreamCreateWithPriority(&c->edge_stream, cudaStreamNonBlocking, prio_hi)) != cudaSuccess) return bail(e, "cudaStreamCreate");
    cudaEventCreate(&c->ev_t0); cudaEventCreate(&c->ev_t1);
    cudaEventCreateWithFlags(&c->ev_bnd, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_comm, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_int, cudaEventDisableTiming);
    cudaEventCreateWithFlags(&c->ev_p3, cudaEventDisableTiming);
    for (int s = 0; s < 2; ++s) {
        cudaEventCreateWithFlags(&c->ev_packed[s], cudaEventDisableTiming);
        cudaEventCreateWithFlags(&c->ev_free[s], cudaEventDisableTiming);
    }
    *out = c;
    return 0;
}

int fdtd_destroy(fdtd_handle c) {
    if (!c) return 0;
    cudaSetDevice(c->cfg.device);
    cudaDeviceSynchronize();
    if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
    if (c->probe_dev) cudaFree(c->probe_dev);
    if (c->stage) cudaFree(c->stage);
    resident_free_tables(c);
    for (FrameMaps& m : c->fm) m.release();
    for (cudaEvent_t e : c->kt_ev) cudaEventDestroy(e);
    for (int s = 0; s < 2; ++s) {
        if (c->snap_dev[s]) cudaFree(c->snap_dev[s]);
        if (c->snap_pin[s]) cudaFreeHost(c->snap_pin[s]);
        if (c->ev_packed[s]) cudaEventDestroy(c->ev_packed[s]);
        if (c->ev_free[s]) cudaEventDestroy(c->ev_free[s]);
        if (c->view[s].dev) cudaFree(c->view[s].dev);
        if (c->view[s].pin) cudaFreeHost(c->view[s].pin);
        if (c->view[s].done) cudaEventDestroy(c->view[s].done);
    }
    for (auto& t : c->split_tiles) if (t.dev) cudaFree(t.dev);
    if (c->scan_dev) cudaFree(c->scan_dev);
    for (cudaEvent_t e : c->band_up) cudaEventDestroy(e);
    for (cudaEvent_t e : c->band_done) cudaEventDestroy(e);
    for (cudaEvent_t e : c->band_edge) cudaEventDestroy(e);
    if (c->d2h_stream) cudaStreamDestroy(c->d2h_stream);
    if (c->edge_stream2) cudaStreamDestroy(c->edge_stream2);
    for (cudaEvent_t e : {c->ev_t0, c->ev_t1, c->ev_bnd, c->ev_comm, c->ev_int, c->ev_p3}) if (e) cudaEventDestroy(e);
    for (cudaStream_t s : {c->stream, c->copy_stream, c->comm_stream, c->edge_stream}) if (s) cudaStreamDestroy(s);
    delete c;
    return 0;
}

int fdtd_bind_buffers(fdtd_handle c, void* h0, void* h1, void* ex0, void* ex1, void* ey0, void* ey1,
                      void* eps_coef, void* mu_coef) {
    if (!c) return FDTD_ERR_ARG;
    void* all[8] = {h0, h1, ex0, ex1, ey0, ey1, eps_coef, mu_coef};
    for (void* p : all) {
        if (!p) return fail(c, FDTD_ERR_ARG, "null device buffer");
        if (reinterpret_cast<uintptr_t>(p) % 128) return fail(c, FDTD_ERR_ARG, "device buffers must be 128-byte aligned");
    }
    c->gen[0][0] = h0; c->gen[1][0] = h1; c->gen[0][1] = ex0; c->gen[1][1] = ex1; c->gen[0][2] = ey0; c->gen[1][2] = ey1;
    c->epc = eps_coef; c->muc = mu_coef;
    c->bound = true; c->cur = 0;
    c->mu_uniform = false;                                       // nothing is known about freshly bound planes
    return 0;
}

static int plan_frame_export(const fdtd_config* cfg, int32_t steps_per_pass, int32_t n_boxes, const int64_t* boxes,
                             fdtd_frame_plan* plan, uint8_t* skip, uint8_t* masks) {
    fdtd_ctx* c = nullptr;
    if (!cfg || !plan || steps_per_pass < 2 || steps_per_pass > 8 || n_boxes < 0 || (n_boxes > 0 && !boxes))
        return fail(c, FDTD_ERR_ARG, "bad argument");
    if (cfg->rows < 3 || cfg->cols < 3 || cfg->tile_rows <= 0 || cfg->block_warps <= 0 || cfg->row_begin < 0 ||
        cfg->row_end > cfg->rows || cfg->row_end <= cfg->row_begin)
        return fail(c, FDTD_ERR_ARG, "bad geometry");
    std::vector<FrameBox> mods;
    for (int k = 0; k < n_boxes; ++k) mods.push_back({boxes[4 * k], boxes[4 * k + 1], boxes[4 * k + 2], boxes[4 * k + 3]});
    const int V2 = cfg->dtype != FDTD_F32 ? 2 : 4, K = steps_per_pass;
    FrameMaps m;
    plan_frame(FrameGeom{cfg->rows, cfg->cols, cfg->row_begin, cfg->row_end, cfg->tile_rows, cfg->block_warps, K, V2, 0, 0},
               mods, m);
    const int margin = (K + V2 - 1) / V2;
    plan->k_tile_rows = cfg->tile_rows; plan->k_tile_cols = (32 - 2 * margin) * V2; plan->k_load_cols = 32 * V2;
    plan->k_margin_cols = margin * V2; plan->k_tiles_y = m.nty; plan->k_tiles_x = m.ntx2;
    plan->k_row_lo = m.k_row_lo; plan->k_row_hi = m.k_row_hi;
    plan->frame_tile_rows = m.frame_tile_rows; plan->frame_tile_cols = 32 * m.frame_vec;
    plan->frame_tiles_y = m.frame_nty; plan->frame_tiles_x = m.ntx1;
    if (skip && m.nty > 0) std::memcpy(skip, m.h_skip.data(), (size_t)m.nty * m.ntx2);
    if (masks)
        for (int d = 0; d < K; ++d)
            std::memcpy(masks + (size_t)d * m.frame_nty * m.ntx1, m.h_mask[d].data(), (size_t)m.frame_nty * m.ntx1);
    return 0;
}

int fdtd_plan_frame(const fdtd_config* cfg, int32_t steps_per_pass, int32_t n_boxes, const int64_t* boxes,
                    fdtd_frame_plan* plan, uint8_t* skip, uint8_t* masks) {
    return plan_frame_export(cfg, steps_per_pass, n_boxes, boxes, plan, skip, masks);
}

int fdtd_plan_edge(const fdtd_config* cfg, int32_t steps_per_pass, int32_t halo_depth, int32_t n_boxes, const int64_t* boxes,
                   const int32_t* box_is_pulse, fdtd_frame_plan* plan, uint8_t* skip, int32_t* n_edge_tiles,
                   int32_t* edge_tiles, int32_t max_edge_tiles) {
    fdtd_ctx* c = nullptr;
    if (!cfg || !plan || steps_per_pass < 2 || steps_per_pass > 8 || n_boxes < 0 || (n_boxes > 0 && !boxes) ||
        halo_depth < 1 || halo_depth > 9)
        return fail(c, FDTD_ERR_ARG, "bad argument");
    if (cfg->rows < 3 || cfg->cols < 3 || cfg->tile_rows <= 0 || cfg->row_begin < 0 || cfg->row_end > cfg->rows ||
        cfg->row_end <= cfg->row_begin)
        return fail(c, FDTD_ERR_ARG, "bad geometry");
    const bool whole = cfg->row_begin == 0 && cfg->row_end == cfg->rows;
    if (!whole && halo_depth < steps_per_pass + 1)
        return fail(c, FDTD_ERR_ARG, "a y-slab needs halos of steps_per_pass + 1 rows for the edge plan");
    std::vector<FrameBox> mods;
    for (int k = 0; k < n_boxes; ++k)
        mods.push_back({boxes[4 * k], boxes[4 * k + 1], boxes[4 * k + 2], boxes[4 * k + 3],
                        !box_is_pulse ? (fdtd::EDGE_SRC | fdtd::EDGE_MISC) : (box_is_pulse[k] ? fdtd::EDGE_SRC : fdtd::EDGE_MISC)});
    const int V2 = cfg->dtype != FDTD_F32 ? 2 : 4, K = steps_per_pass;
    FrameMaps m;
    const int64_t ht = (halo_depth > 1 && cfg->row_begin > 0) ? halo_depth : 0;
    const int64_t hb = (halo_depth > 1 && cfg->row_end < cfg->rows) ? halo_depth : 0;
    plan_edge(FrameGeom{cfg->rows, cfg->cols, cfg->row_begin, cfg->row_end, cfg->tile_rows, std::max(cfg->block_warps, 1), K, V2,
                        ht, hb, true}, mods, m);
    const int margin = (K + V2 - 1) / V2;
    plan->k_tile_rows = cfg->tile_rows; plan->k_tile_cols = (32 - 2 * margin) * V2; plan->k_load_cols = 32 * V2;
    plan->k_margin_cols = margin * V2; plan->k_tiles_y = m.nty; plan->k_tiles_x = m.ntx2;
    plan->k_row_lo = m.k_row_lo; plan->k_row_hi = m.k_row_hi;
    plan->frame_tile_rows = 0; plan->frame_tile_cols = 0; plan->frame_tiles_y = 0; plan->frame_tiles_x = 0;
    if (skip && m.nty > 0) std::memcpy(skip, m.h_skip.data(), (size_t)m.nty * m.ntx2);
    if (n_edge_tiles) *n_edge_tiles = (int32_t)m.h_edge.size();
    if (edge_tiles)
        for (size_t k = 0; k < m.h_edge.size() && (int32_t)k < max_edge_tiles; ++k) {
            edge_tiles[4 * k] = m.h_edge[k].tx; edge_tiles[4 * k + 1] = m.h_edge[k].ta; edge_tiles[4 * k + 2] = m.h_edge[k].tb;
            edge_tiles[4 * k + 3] = ((int)k < m.n_edge_near ? 1 : 0) | (m.h_edge[k].pad << 8);
        }
    return 0;
}

int fdtd_bind_scratch(fdtd_handle c, void* h2, void* ex2, void* ey2) {
    if (!c) return FDTD_ERR_ARG;
    for (void* p : {h2, ex2, ey2}) {
        if (!p) return fail(c, FDTD_ERR_ARG, "null device buffer");
        if (reinterpret_cast<uintptr_t>(p) % 128) return fail(c, FDTD_ERR_ARG, "device buffers must be 128-byte aligned");
    }
    if (!c->bound) return fail(c, FDTD_ERR_STATE, "bind the two field generations first");
    if (c->ngen >= 4) return fail(c, FDTD_ERR_STATE, "at most two scratch generations");
    c->gen[c->ngen][0] = h2; c->gen[c->ngen][1] = ex2; c->gen[c->ngen][2] = ey2;
    c->ngen++;
    return 0;
}

// Launch geometry the library picks when the caller left it open (cfg.tile_rows / block_warps <= 0 at fdtd_create).
static void auto_geometry(fdtd_ctx* c) {
    const int steps_per_pass = c->temporal;
    // measured on the 8192 x 65536 slab (profiles/sweep_r1*.jsonl, r2_sweep_*): the K-step kernel wants taller tiles as K
    // grows (its prologue is 2K-1 rows) and one warp per CTA; with the edge tiles cut to their own short height (frame
    // mode 3) K = 7, 8 like them much taller still (K = 8: 56 rows 552, 128 rows 607, 512 rows 632 Gcell/s)
    static const int rows_for_k[9] = {6, 6, 8, 12, 16, 32, 48, 48, 56};
    static const int rows_for_k_edge[9] = {6, 6, 8, 12, 16, 32, 48, 256, 512};
    if (c->auto_tile_rows) {
        int tr = rows_for_k[steps_per_pass];
        if (c->frame_mode == 3 && rows_for_k_edge[steps_per_pass] > tr) {
            // ... as long as the plain tiles still fill the GPU several times over (8 single-warp CTAs per SM at 250
            // registers): a grid much smaller than the 8192 x 65536 slab keeps shorter tiles
            const int V2 = c->cfg.dtype != FDTD_F32 ? 2 : 4, K = steps_per_pass;
            const int64_t out_cols = (32 - 2 * ((K + V2 - 1) / V2)) * V2;
            const int64_t ntx = (c->cfg.cols + 1 + out_cols - 1) / out_cols;
            const int64_t want_nty = (4 * 8 * 148 + ntx - 1) / ntx;
            const int64_t fit = (c->owned / std::max<int64_t>(want_nty, 1)) / 8 * 8;
            tr = (int)std::min<int64_t>(rows_for_k_edge[steps_per_pass], std::max<int64_t>(tr, fit));
        }
        c->cfg.tile_rows = tr;
    }
    if (c->auto_warps) c->cfg.block_warps = steps_per_pass == 1 ? 4 : (steps_per_pass == 4 ? 2 : 1);
}

int fdtd_set_temporal_blocking(fdtd_handle c, int32_t steps_per_pass) {
    if (!c) return FDTD_ERR_ARG;
    if (steps_per_pass < 1 || steps_per_pass > 8) return fail(c, FDTD_ERR_ARG, "steps_per_pass must be 1..8");
    // Scratch generations (fdtd_bind_scratch) are needed only by the masked pass chain (frame mode 0, one-row halos, a
    // TF/SF "right" line at column 0).  Frame mode 3 reads generation A and writes generation B and nothing else; a run
    // that finds neither stepK_edge usable nor the scratch bound steps call by call (fdtd_run).
    c->temporal = steps_per_pass;
    auto_geometry(c);
    for (FrameMaps& m : c->fm) m.key.clear();
    return 0;
}

int fdtd_set_halo_depth(fdtd_handle c, int32_t depth) {
    if (!c) return FDTD_ERR_ARG;
    if (c->bound) return fail(c, FDTD_ERR_STATE, "set the halo depth before fdtd_bind_buffers (it sizes the arrays)");
    if (depth < 1 || depth > 9) return fail(c, FDTD_ERR_ARG, "halo depth must be 1..9");
    const bool up = c->cfg.row_begin > 0, down = c->cfg.row_end < c->cfg.rows;
    if (depth > 1 && (up || down) && c->owned < depth)
        return fail(c, FDTD_ERR_ARG, "a slab must own at least as many rows as its halos are deep");
    c->halo_depth = depth;
    c->halo_top = up ? depth : 0;
    c->halo_bot = (depth > 1 && down) ? depth : 0;
    c->row_off = c->cfg.row_begin - c->halo_top;
    c->lrows = fdtd_local_rows_halo(&c->cfg, depth);
    for (FrameMaps& m : c->fm) m.key.clear();
    return 0;
}

int fdtd_set_resident(fdtd_handle c, int32_t enable, int64_t max_cells) {
    if (!c) return FDTD_ERR_ARG;
    if (max_cells < 0) return fail(c, FDTD_ERR_ARG, "max_cells must be >= 0 (0 keeps the default)");
    c->resident = enable ? 1 : 0;
    if (max_cells > 0) c->resident_max_cells = max_cells;
    return 0;
}

int fdtd_set_kstep_variant(fdtd_handle c, int32_t variant) {
    if (!c) return FDTD_ERR_ARG;
    if (variant < 0 || variant > 1) return fail(c, FDTD_ERR_ARG, "K-step kernel variant must be 0 or 1");
    c->kstep_variant = variant;
    return 0;
}

int fdtd_set_frame_mode(fdtd_handle c, int32_t mode) {
    if (!c) return FDTD_ERR_ARG;
    if (mode != 0 && mode != 3) return fail(c, FDTD_ERR_ARG, "frame mode must be 0 (masked pass chain) or 3 (stepK_edge)");
    c->frame_mode = mode;
    auto_geometry(c);
    for (FrameMaps& m : c->fm) m.key.clear();
    return 0;
}

int fdtd_resident_active(fdtd_handle c) { return c ? (resident_eligible(c) ? 1 : 0) : FDTD_ERR_ARG; }

#define NEED_BOUND() do { if (!c) return FDTD_ERR_ARG; if (!c->bound) return fail(c, FDTD_ERR_STATE, "buffers not bound"); \
    CK(cudaSetDevice(c->cfg.device)); } while (0)

int fdtd_set_host_row_origin(fdtd_handle c, int64_t first_global_row) {
    if (!c) return FDTD_ERR_ARG;
    if (first_global_row < 0 || first_global_row > c->row_off) return fail(c, FDTD_ERR_ARG, "host arrays must start at or above the slab's first stored row");
    c->host_row0 = first_global_row;
    return 0;
}

int fdtd_upload_coeffs(fdtd_handle c, const double* eps_coef, const double* mu_coef, int64_t host_ld) {
    Range nvtx("fdtd_upload_coeffs");
    NEED_BOUND();
    if (!eps_coef || !mu_coef || host_ld < c->cfg.cols) return fail(c, FDTD_ERR_ARG, "bad coefficient arrays");
    const int64_t g0 = c->row_off, n = c->cfg.row_end + c->halo_bot - c->row_off;     // every stored h-row
    int rc;
    if (c->cfg.dtype != FDTD_F32) {                      // float64 coefficient planes (also under split field storage)
        if ((rc = h2d_plane<double>(c, c->epc, eps_coef, host_ld, g0, n, c->cfg.cols))) return rc;
        if ((rc = h2d_plane<double>(c, c->muc, mu_coef, host_ld, g0, n, c->cfg.cols))) return rc;
    } else {
        if ((rc = h2d_plane<float>(c, c->epc, eps_coef, host_ld, g0, n, c->cfg.cols))) return rc;
        if ((rc = h2d_plane<float>(c, c->muc, mu_coef, host_ld, g0, n, c->cfg.cols))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    return scan_mu_uniform(c);
}

int fdtd_fill_synthetic_coeffs(fdtd_handle c, uint64_t seed, double dt_over_ds, double epsilon0, double mu0) {
    NEED_BOUND();
    dim3 grid((unsigned)((c->cfg.cols + 255) / 256), (unsigned)std::min<int64_t>(c->lrows, 65535));
    if (c->cfg.dtype != FDTD_F32)
        CK(fdtd::launch_synth_coeffs<double>(grid, c->stream, (double*)c->epc, (double*)c->muc, c->pitch, (int)c->lrows,
            (int)c->cfg.cols, (int)c->row_off, (int)c->cfg.rows, c->cfg.cols, seed, dt_over_ds, epsilon0, mu0));
    else
        CK(fdtd::launch_synth_coeffs<float>(grid, c->stream, (float*)c->epc, (float*)c->muc, c->pitch, (int)c->lrows,
            (int)c->cfg.cols, (int)c->row_off, (int)c->cfg.rows, c->cfg.cols, seed, dt_over_ds, epsilon0, mu0));
    c->stats.kernel_launches++;
    CK(cudaStreamSynchronize(c->stream));
    return scan_mu_uniform(c);
}

int fdtd_paint_coeffs(fdtd_handle c, int32_t n, const fdtd_shape* shapes, double eps_background, double mu_background,
                      double dt_over_ds) {
    NEED_BOUND();
    if (n < 0 || (n > 0 && !shapes)) return fail(c, FDTD_ERR_ARG, "bad shape list");
    std::vector<fdtd::ShapeDev> host((size_t)std::max(n, 1));
    for (int k = 0; k < n; ++k) {
        if (shapes[k].kind < FDTD_SHAPE_RECT || shapes[k].kind > FDTD_SHAPE_CONVEX) return fail(c, FDTD_ERR_ARG, "bad shape kind");
        host[k].kind = shapes[k].kind; host[k].pad = 0;
        for (int q = 0; q < 6; ++q) host[k].a[q] = shapes[k].a[q];
        host[k].eps = shapes[k].eps; host[k].mu = shapes[k].mu;
    }
    fdtd::ShapeDev* dev = nullptr;
    CK(cudaMalloc(&dev, host.size() * sizeof(fdtd::ShapeDev)));
    CK(cudaMemcpyAsync(dev, host.data(), host.size() * sizeof(fdtd::ShapeDev), cudaMemcpyHostToDevice, c->stream));
    dim3 grid((unsigned)((c->cfg.cols + 255) / 256), (unsigned)std::min<int64_t>(c->lrows, 65535));
    cudaError_t e;
    if (c->cfg.dtype != FDTD_F32)
        e = fdtd::launch_paint_coeffs<double>(grid, c->stream, (double*)c->epc, (double*)c->muc, c->pitch, (int)c->lrows,
            (int)c->cfg.cols, (int)c->row_off, (int)c->cfg.rows, dev, n, eps_background, mu_background, dt_over_ds);
    else
        e = fdtd::launch_paint_coeffs<float>(grid, c->stream, (float*)c->epc, (float*)c->muc, c->pitch, (int)c->lrows,
            (int)c->cfg.cols, (int)c->row_off, (int)c->cfg.rows, dev, n, eps_background, mu_background, dt_over_ds);
    c->stats.kernel_launches++;
    cudaError_t e2 = cudaStreamSynchronize(c->stream);
    cudaFree(dev);
    if (e != cudaSuccess || e2 != cudaSuccess)
        return fail(c, FDTD_ERR_CUDA, std::string("paint_coeffs: ") + cudaGetErrorString(e != cudaSuccess ? e : e2));
    return scan_mu_uniform(c);
}

// The library remembers one property of the coefficient planes (a uniform mu_coef plane lets the K-step kernel skip that
// stream).  It is re-established by every entry point that writes the planes; a caller that writes the caller-owned
// tensors directly says so here.
int fdtd_rescan_coeffs(fdtd_handle c) {
    NEED_BOUND();
    if (int rc = quiesce(c)) return rc;
    return scan_mu_uniform(c);
}
int fdtd_uniform_mu(fdtd_handle c) { return c ? (c->mu_uniform ? 1 : 0) : FDTD_ERR_ARG; }

// Host-side accessors must not race with work still in flight on the side streams (a halo block of the last exchange
// landing after an upload, a frame / edge launch still writing): they wait for all of them first.
static int quiesce(fdtd_ctx* c) {
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->edge_stream));
    CK(cudaStreamSynchronize(c->comm_stream));
    return 0;
}

int fdtd_zero_fields(fdtd_handle c) {
    NEED_BOUND();
    if (int rc = quiesce(c)) return rc;
    for (int g = 0; g < c->ngen; ++g)
        for (int f = 0; f < 3; ++f)
            CK(cudaMemsetAsync(c->gen[g][f], 0, (size_t)c->lrows * c->pitch * c->esize, c->stream));
    CK(cudaStreamSynchronize(c->stream));
    c->cur = 0;
    return 0;
}

int fdtd_set_fields(fdtd_handle c, const double* hh, const double* ex, const double* ey) {
    Range nvtx("fdtd_set_fields");
    NEED_BOUND();
    if (int rc = quiesce(c)) return rc;
    const int64_t g0 = c->row_off, n = c->cfg.row_end + c->halo_bot - c->row_off, NC = c->cfg.cols;
    const int a = c->cur;
    int rc;
    if (c->split) {
        if ((rc = split_h2d_field(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = split_h2d_field(c, c->gen[a][1], ex, NC, g0, n + 1, NC))) return rc;
        if ((rc = split_h2d_field(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    } else if (c->cfg.dtype == FDTD_F64) {
        if ((rc = h2d_plane<double>(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = h2d_plane<double>(c, c->gen[a][1], ex, NC, g0, n + 1, NC))) return rc;
        if ((rc = h2d_plane<double>(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    } else {
        if ((rc = h2d_plane<float>(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = h2d_plane<float>(c, c->gen[a][1], ex, NC, g0, n + 1, NC))) return rc;
        if ((rc = h2d_plane<float>(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fdtd_get_fields(fdtd_handle c, double* hh, double* ex, double* ey) {
    Range nvtx("fdtd_get_fields");
    NEED_BOUND();
    if (int rc = quiesce(c)) return rc;
    const int64_t g0 = c->cfg.row_begin, n = c->owned, NC = c->cfg.cols;
    const int64_t nex = n + (c->cfg.row_end == c->cfg.rows ? 1 : 0);
    const int a = c->cur;
    int rc;
    if (c->split) {
        if ((rc = split_d2h_field(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = split_d2h_field(c, c->gen[a][1], ex, NC, g0, nex, NC))) return rc;
        if ((rc = split_d2h_field(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    } else if (c->cfg.dtype == FDTD_F64) {
        if ((rc = d2h_plane<double>(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = d2h_plane<double>(c, c->gen[a][1], ex, NC, g0, nex, NC))) return rc;
        if ((rc = d2h_plane<double>(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    } else {
        if ((rc = d2h_plane<float>(c, c->gen[a][0], hh, NC, g0, n, NC))) return rc;
        if ((rc = d2h_plane<float>(c, c->gen[a][1], ex, NC, g0, nex, NC))) return rc;
        if ((rc = d2h_plane<float>(c, c->gen[a][2], ey, NC + 1, g0, n, NC + 1))) return rc;
    }
    CK(cudaStreamSynchronize(c->stream));
    return 0;
}

int fdtd_set_sources(fdtd_handle c, int32_t n, const fdtd_source* src, int64_t table_len,
                     const double* mag, const double* mag_z, const uint8_t* active, const double* last_mag) {
    if (!c) return FDTD_ERR_ARG;
    for (FrameMaps& m : c->fm) m.key.clear();            // temporal-blocking frame maps depend on the scene
    c->rs_valid = false;                                  // so do the resident kernel's per-call pulse records
    if (n < 0 || n > 64) return fail(c, FDTD_ERR_ARG, "at most 64 pulses");
    if (n > 0 && (!src || !mag || !mag_z || !active || !last_mag || table_len <= 0))
        return fail(c, FDTD_ERR_ARG, "null source table");
    for (int p = 0; p < n; ++p) {
        const fdtd_source& s = src[p];
        const bool bad_kind = s.kind < FDTD_SRC_POINT || s.kind > FDTD_SRC_DOWN;
        bool bad = bad_kind || s.row < 0 || s.col < 0 || s.row >= c->cfg.rows || s.col >= c->cfg.cols;
        if (s.kind == FDTD_SRC_UP && s.row + 1 >= c->cfg.rows) bad = true;    // h[1+row,:] must exist (solver.py:198)
        if (bad) return fail(c, FDTD_ERR_ARG, "pulse " + std::to_string(p) + " lies outside the grid");
    }
    c->n_src = n; c->table_len = n ? table_len : 0;
    for (int p = 0; p < n; ++p) c->src[p] = src[p];
    const size_t tot = (size_t)n * (size_t)c->table_len;
    c->mag.assign(mag, mag + tot); c->magz.assign(mag_z, mag_z + tot);
    c->active.assign(active, active + tot);
    c->last.assign(last_mag, last_mag + (n ? table_len : 0));
    return 0;
}

int fdtd_set_reflectors(fdtd_handle c, int32_t n, const int64_t* rects) {
    if (!c) return FDTD_ERR_ARG;
    for (FrameMaps& m : c->fm) m.key.clear();
    if (n < 0 || n > FDTD_MAX_RECT) return fail(c, FDTD_ERR_ARG, "at most 24 reflector rectangles");
    if (n > 0 && !rects) return fail(c, FDTD_ERR_ARG, "null rectangle list");
    c->n_rect = n;
    const int64_t NR = c->cfg.rows, NC = c->cfg.cols;
    for (int r = 0; r < n; ++r) {
        c->rh[r] = norm_rect(rects + 4 * r, NR, NC);
        c->rex[r] = norm_rect(rects + 4 * r, NR + 1, NC);
        c->rey[r] = norm_rect(rects + 4 * r, NR, NC + 1);
    }
    return 0;
}

int fdtd_set_probes(fdtd_handle c, int32_t n, const int64_t* ij) {
    if (!c) return FDTD_ERR_ARG;
    for (FrameMaps& m : c->fm) m.key.clear();
    if (n < 0 || n > FDTD_MAX_PROBE) return fail(c, FDTD_ERR_ARG, "at most 16 probes");
    if (n > 0 && !ij) return fail(c, FDTD_ERR_ARG, "null probe list");
    cudaSetDevice(c->cfg.device);
    if (c->probe_dev) { cudaStreamSynchronize(c->stream); cudaFree(c->probe_dev); c->probe_dev = nullptr; }
    c->probe_cap = 0; c->probe_count = 0;
    for (int k = 0; k < n; ++k) {
        if (ij[2 * k] < 0 || ij[2 * k] >= c->cfg.rows || ij[2 * k + 1] < 0 || ij[2 * k + 1] >= c->cfg.cols)
            return fail(c, FDTD_ERR_ARG, "probe outside the grid");
        c->probe_i[k] = (int)ij[2 * k]; c->probe_j[k] = (int)ij[2 * k + 1];
    }
    c->n_probe = n;
    return 0;
}

int fdtd_probe_fetch(fdtd_handle c, double* out, int64_t max_records, int64_t* n_records) {
    if (!c) return FDTD_ERR_ARG;
    CK(cudaSetDevice(c->cfg.device));
    const int64_t n = std::min(max_records, c->probe_count);
    if (n_records) *n_records = c->probe_count;
    if (n > 0 && out && c->n_probe) {
        CK(cudaStreamSynchronize(c->stream));
        CK(cudaMemcpy(out, c->probe_dev, (size_t)n * c->n_probe * 3 * sizeof(double), cudaMemcpyDeviceToHost));
    }
    return 0;
}

int fdtd_probe_clear(fdtd_handle c) {
    if (!c) return FDTD_ERR_ARG;
    c->probe_count = 0;
    return 0;
}

int fdtd_step(fdtd_handle c, int64_t step) {
    NEED_BOUND();
    if (resident_eligible(c)) {                          // a one-call launch of step_resident: no slow corner CTA
        if (int rc = ensure_probe_capacity(c, 1)) return rc;
        return resident_any(c, step, 1);
    }
    return step_any(c, step);
}

// The run loops of solver.py:288-292 / :331-337.  wait = false (fdtd_run_enqueue): everything is queued on the context's
// streams and the host returns at once -- consecutive runs chain on the device exactly like the groups of one run do.
static int run_impl(fdtd_handle c, int64_t first_step, int64_t n_steps, int64_t snapshot_every,
                    double* snapshots_host, int64_t max_snapshots, int64_t total_calls, int64_t* n_snapshots, bool wait) {
    NEED_BOUND();
    Range nvtx(wait ? "fdtd_run" : "fdtd_run_enqueue");
    if (n_snapshots) *n_snapshots = 0;
    if (n_steps < 0) return fail(c, FDTD_ERR_ARG, "negative step count");
    const bool snaps = snapshot_every > 0 && snapshots_host != nullptr;
    if (snaps) {
        for (int s = 0; s < 2; ++s) {                   // each checked on its own: a failed half is retried next call
            if (!c->snap_dev[s]) CK(cudaMalloc(&c->snap_dev[s], (size_t)c->owned * c->cfg.cols * 8));
            if (!c->snap_pin[s]) CK(cudaHostAlloc(&c->snap_pin[s], (size_t)c->owned * c->cfg.cols * 8, cudaHostAllocDefault));
        }
    }
    if (int rc = ensure_probe_capacity(c, n_steps)) return rc;
    int64_t k = 0;
    int rc = 0;
    const bool resident = resident_eligible(c);
    // steps per group: what the caller asked for, what the bound scratch generations allow, and -- on a slab with deep
    // halos -- what the halos cover (a group of K steps reads K+1 rows of the neighbours)
    int kmax_pre = (c->cfg.kernel == FDTD_KERNEL_STREAM) ? c->temporal : 1;
    if (c->halo_depth > 1 && !(c->cfg.row_begin == 0 && c->cfg.row_end == c->cfg.rows)) kmax_pre = std::min(kmax_pre, c->halo_depth - 1);
    if (kmax_pre >= 2 && !edge_usable(c, kmax_pre))      // only the pass chain needs the scratch generations
        kmax_pre = std::min(kmax_pre, c->ngen >= 4 ? 8 : (c->ngen == 3 ? 2 : 1));
    if (kmax_pre >= 2 && !resident && !(c->split && (c->split_all_edge || !edge_usable(c, kmax_pre)))) {
        // Build the frame maps of every group size this run will use BEFORE the timed region (host work + a sync each).
        bool seen[9] = {};
        for (int64_t n = first_step; n < first_step + n_steps;) {
            int64_t g = std::min<int64_t>(kmax_pre, first_step + n_steps - n);
            if (snaps) g = std::min<int64_t>(g, snapshot_every - (n % snapshot_every));
            if (c->n_src > 0) g = std::min<int64_t>(g, c->table_len - n);
            if (g >= 2 && !seen[g]) {
                seen[g] = true;
                const int rc0 = c->cfg.dtype != FDTD_F32 ? build_frame_maps<double>(c, n, (int)g) : build_frame_maps<float>(c, n, (int)g);
                if (rc0) return rc0;
            }
            n += std::max<int64_t>(g, 1);
        }
        if (c->frame_mode != 0)                          // stepK_edge reads the pulse records of resident mode
            if (int rc0 = c->cfg.dtype != FDTD_F32 ? resident_tables<double>(c) : resident_tables<float>(c)) return rc0;
    }
    if (resident)                                        // upload the pulse records before the timed region, too
        if (int rc0 = c->cfg.dtype == FDTD_F64 ? resident_tables<double>(c) : resident_tables<float>(c)) return rc0;
    CK(cudaEventRecord(c->ev_t0, c->stream));
    const int kmax = kmax_pre;
    for (int64_t n = first_step; n < first_step + n_steps; ++n) {
        // largest group that ends no later than the run, the next snapshot and the source tables
        int64_t g = std::min<int64_t>(kmax, first_step + n_steps - n);
        if (snaps) g = std::min<int64_t>(g, snapshot_every - (n % snapshot_every));
        if (c->n_src > 0) g = std::min<int64_t>(g, c->table_len - n);
        if (resident) {
            // every call up to the next snapshot (or the end of the run / of the pulse tables) in one cooperative launch
            int64_t r = std::min<int64_t>(first_step + n_steps - n, 1 << 20);
            if (snaps) r = std::min<int64_t>(r, snapshot_every - (n % snapshot_every));
            if (c->n_src > 0) r = std::min<int64_t>(r, c->table_len - n);
            if (r >= 1) {
                if ((rc = resident_any(c, n, r))) break;
                n += r - 1;
            } else if ((rc = step_any(c, n))) break;     // past the pulse tables: reports the range error
        } else if (g >= 2) {
            if ((rc = group_any(c, n, (int)g))) break;
            n += g - 1;                                  // g steps done
        } else if ((rc = step_any(c, n))) break;
        if (snaps && (n + 1) % snapshot_every == 0) {
            if (k >= max_snapshots) { rc = fail(c, FDTD_ERR_ARG, "snapshot buffer too small"); break; }
            double* dst = snapshots_host + (size_t)k * c->owned * c->cfg.cols;
            rc = c->cfg.dtype != FDTD_F32 ? snapshot_impl<double>(c, (int)(k & 1), dst, n + 1, total_calls)
                                          : snapshot_impl<float>(c, (int)(k & 1), dst, n + 1, total_calls);
            if (rc) break;
            ++k;
        }
    }
    if (c->comm) cudaStreamWaitEvent(c->stream, c->ev_comm, 0);      // the timed region ends after the last halo landed
    cudaEventRecord(c->ev_t1, c->stream);
    if (!wait) {
        if (n_snapshots) *n_snapshots = k;
        if (rc) return rc;
        CK(cudaGetLastError());
        return 0;
    }
    cudaError_t e1 = cudaStreamSynchronize(c->stream);
    cudaError_t e2 = cudaStreamSynchronize(c->copy_stream);
    cudaError_t e3 = cudaStreamSynchronize(c->comm_stream);
    if (e3 == cudaSuccess) e3 = cudaStreamSynchronize(c->edge_stream);
    if (n_snapshots) *n_snapshots = k;
    if (rc) return rc;
    for (cudaError_t e : {e1, e2, e3})
        if (e != cudaSuccess) return fail(c, FDTD_ERR_CUDA, std::string("fdtd_run: ") + cudaGetErrorString(e));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
    c->stats.last_run_ms = ms;
    return 0;
}

int fdtd_run(fdtd_handle c, int64_t first_step, int64_t n_steps, int64_t snapshot_every,
             double* snapshots_host, int64_t max_snapshots, int64_t total_calls, int64_t* n_snapshots) {
    return run_impl(c, first_step, n_steps, snapshot_every, snapshots_host, max_snapshots, total_calls, n_snapshots, true);
}

int fdtd_run_enqueue(fdtd_handle c, int64_t first_step, int64_t n_steps) {
    return run_impl(c, first_step, n_steps, 0, nullptr, 0, 0, nullptr, false);
}

}  // extern "C"

// ---- streamed solve: upload -> calls -> download as one pipeline over row bands (end-to-end path) ------------
// The band plan of a group size: tile rows and edge tiles grouped by band = first row / band_rows.
template <typename T>
static int build_band_plan(fdtd_ctx* c, int K, int64_t band_rows, int n_bands) {
    FrameMaps& m = c->fm[K];
    if (m.band_rows == band_rows && (int)m.band_ty0.size() == n_bands) return 0;
    auto band_of = [&](int64_t row) { return (int)std::min<int64_t>(std::max<int64_t>(row, 0) / band_rows, n_bands - 1); };
    m.band_ty0.assign(n_bands, 0); m.band_ny.assign(n_bands, 0);
    for (int ty = m.nty - 1; ty >= 0; --ty) {
        const int b = band_of(m.k_row_lo + (int64_t)ty * c->cfg.tile_rows);
        m.band_ty0[b] = ty; m.band_ny[b]++;
    }
    std::vector<fdtd::EdgeTile> sorted = m.h_edge;
    std::stable_sort(sorted.begin(), sorted.end(), [&](const fdtd::EdgeTile& x, const fdtd::EdgeTile& y) {
        const int bx = band_of(x.ta), by = band_of(y.ta);
        return bx != by ? bx < by : x.pad < y.pad;
    });
    m.band_runs.assign(n_bands, {});
    for (size_t k = 0; k < sorted.size(); ++k) {
        auto& runs = m.band_runs[band_of(sorted[k].ta)];
        if (runs.empty() || runs.back().flags != sorted[k].pad) runs.push_back({(int)k, 0, sorted[k].pad});
        runs.back().count++;
    }
    if (m.d_band_cap < sorted.size()) {
        if (m.d_band) cudaFree(m.d_band);
        m.d_band = nullptr; m.d_band_cap = 0;
        CK(cudaMalloc(&m.d_band, sorted.size() * sizeof(fdtd::EdgeTile))); m.d_band_cap = sorted.size();
    }
    if (!sorted.empty()) CK(cudaMemcpy(m.d_band, sorted.data(), sorted.size() * sizeof(fdtd::EdgeTile), cudaMemcpyHostToDevice));
    m.band_rows = band_rows;
    return 0;
}

constexpr int NOT_STREAMED = 0x5eed;                             // run_streamed: this run does not fit the pipeline

// calls n .. n+K-1 for the tiles of one row band: generation A -> B.  The plain tiles go to the launch stream, the edge
// tiles to one of two side streams (by band parity); they are tied together by one event per band (band_edge[b] = "the
// edge tiles of the latest group on band b are done"), so the launch stream never waits for the edge tiles of the unit
// it has just queued -- only for those of the previous group on the bands b-1, b, b+1 that the unit reads (or overwrites
// what they read).  probe_base = index of the group's first probe record.
template <typename T>
static int band_group(fdtd_ctx* c, int64_t n, int K, int A, int B, int band, int n_bands, int64_t probe_base) {
    const FrameMaps& m = c->fm[K];
    cudaStream_t es = (band & 1) ? c->edge_stream2 : c->edge_stream;
    const int b0 = std::max(band - 1, 0), b1 = std::min(band + 1, n_bands - 1);
    const bool has_edge = !m.band_runs[band].empty();
    if (has_edge) {
        CK(cudaEventRecord(c->ev_int, c->stream));               // the plain tiles of the previous group
        CK(cudaStreamWaitEvent(es, c->ev_int, 0));
        for (int nb = b0; nb <= b1; ++nb)                        // its edge tiles on the neighbouring bands (other stream)
            if (nb != band) CK(cudaStreamWaitEvent(es, c->band_edge[nb], 0));
        StepParams<T> PE;
        if (int rc = fill_params<T>(c, PE, n, A, B)) return rc;
        PE.n_src = 0; PE.last_mag = T(0);
        PE.row_lo = (int)c->cfg.row_begin;
        PE.row_hi = (int)(c->cfg.row_end + (c->cfg.row_end == c->cfg.rows ? 1 : 0));
        if (c->n_probe) PE.probe_out = c->probe_dev + (size_t)probe_base * c->n_probe * 3;
        fdtd::ResSources<T> R{(const int*)c->rs_count, (const T*)c->rs_last, (const fdtd::SrcStep<T>*)c->rs_src, c->rs_stride};
        for (const FrameMaps::EdgeRun& run : m.band_runs[band]) {
            CK(launch_edge_any<T>(K, PE, run.flags, m.d_band + run.first, run.count, R, n, es));
            c->stats.kernel_launches++;
        }
    }
    if (m.band_ny[band] > 0) {
        for (int nb = b0; nb <= b1; ++nb) CK(cudaStreamWaitEvent(c->stream, c->band_edge[nb], 0));
        StepParams<T> PK;
        if (int rc = fill_params<T>(c, PK, n, A, B)) return rc;
        PK.n_src = 0; PK.n_rect = 0; PK.n_probe = 0; PK.row_lo = (int)m.k_row_lo; PK.row_hi = (int)m.k_row_hi;
        launch_stepK_any<T>(PK, c, K, m.band_ty0[band], m.band_ny[band], c->stream);
        c->stats.kernel_launches++;
        CK(cudaGetLastError());
    }
    if (has_edge) CK(cudaEventRecord(c->band_edge[band], es));
    return 0;
}

// fdtd_upload_coeffs + (fdtd_zero_fields) + fdtd_run + fdtd_get_fields(h) with the three stages overlapped.  The grid is cut
// into row bands; band b's coefficient rows go up on the copy stream, and as soon as band b+1 has landed the first group
// of calls runs on band b, the second group on band b-1, ... (a group on band b needs the previous group on bands b-1,
// b, b+1 -- K-step cones are at most K+1 rows deep -- so groups advance along diagonals of the (band, group) plane, all on
// the launch stream, in an order that is also safe for the two ping-pong generations).  Behind the last group of a band
// its rows of h leave on a second copy stream, while the bands below are still uploading / stepping: PCIe carries both
// directions at once.  Long runs sweep diagonally only through their first and last n_bands - 1 groups; the groups in
// between are whole-grid launches (fdtd_run's).  Bit-identical to the four separate calls.
template <typename T>
static int run_streamed(fdtd_ctx* c, const double* eps, const double* mu, int64_t host_ld, bool zero, int64_t first_step,
                        int64_t n_steps, double* h_out, int64_t out_ld) {
    Range nvtx("fdtd_run_streamed");
    const int64_t NR = c->cfg.rows, NC = c->cfg.cols;
    // the groups of the run, as fdtd_run forms them
    struct Group { int64_t n; int K; };
    std::vector<Group> groups;
    const int kmax = c->temporal;
    for (int64_t n = first_step; n < first_step + n_steps;) {
        int64_t g = std::min<int64_t>(kmax, first_step + n_steps - n);
        if (c->n_src > 0) g = std::min<int64_t>(g, c->table_len - n);
        if (g < 1) return fail(c, FDTD_ERR_RANGE, "step " + std::to_string(n) + " outside the source tables (length " +
                                                      std::to_string(c->table_len) + ")");
        groups.push_back({n, (int)g});
        n += g;
    }
    if (groups.size() >= 2 && groups.back().K == 1) {            // no single step at the end: 8 + 1 -> 7 + 2, 2 + 1 -> 3
        Group& prev = groups[groups.size() - 2];
        if (prev.K >= 3) { prev.K -= 1; groups.back().n -= 1; groups.back().K = 2; }
        else { prev.K += 1; groups.pop_back(); }
    }
    // one tile plan per group size, for the pulses live in ANY group of that size
    std::vector<uint8_t> keys[9];
    for (const Group& g : groups) {
        if (g.K < 2 || !edge_usable(c, g.K)) return NOT_STREAMED;
        const std::vector<uint8_t> key = frame_key(c, g.n, g.K);
        if (keys[g.K].empty()) keys[g.K] = key;
        else for (int p = 0; p < c->n_src; ++p) keys[g.K][p] |= key[p];
    }
    const int TR = c->cfg.tile_rows;
    // Bands: whole tile rows, at least 2 K + 2 rows (a K-step cone reaches into the neighbouring band only), and about
    // `want` of them -- few enough that one band fills the GPU a couple of times over (a band unit is its own launch),
    // enough that the copies come in pieces.  FDTD_STREAM_BANDS overrides the count (measurement hook).
    int want = 16;                                               // (B200, C5 slab, 20 calls: 4 / 8 / 16 bands -> 45 / 53 / 57.5 Gcell/s end to end)
    if (const char* wb = std::getenv("FDTD_STREAM_BANDS")) { const int v = std::atoi(wb); if (v >= 2 && v <= 64) want = v; }
    const int64_t total_ty = (NR + TR - 1) / TR;
    const int64_t H = (int64_t)TR * std::max<int64_t>((2 * 8 + 2 + TR - 1) / TR, (total_ty + want - 1) / want);
    const int NB = (int)((NR + H - 1) / H);
    if (NB < 2 || (c->ngen != 2 && c->cur > 1)) return NOT_STREAMED;
    struct TwoGenerations {                                      // ping-pong between two generations only, also where
        fdtd_ctx* c; int saved;                                  // group_edge picks the next one
        explicit TwoGenerations(fdtd_ctx* c_) : c(c_), saved(c_->ngen) { c->ngen = 2; }
        ~TwoGenerations() { c->ngen = saved; }
    } two(c);
    for (const Group& g : groups) {
        if (int rc = build_frame_maps<T>(c, g.n, g.K, &keys[g.K])) return rc;
        if (int rc = build_band_plan<T>(c, g.K, H, NB)) return rc;
    }
    struct PinnedMaps {
        fdtd_ctx* c;
        explicit PinnedMaps(fdtd_ctx* c_) : c(c_) { c->maps_pinned = true; }
        ~PinnedMaps() { c->maps_pinned = false; }
    } pinned(c);
    if (int rc = resident_tables<T>(c)) return rc;
    if (int rc = ensure_probe_capacity(c, n_steps)) return rc;
    if (!c->d2h_stream) CK(cudaStreamCreateWithFlags(&c->d2h_stream, cudaStreamNonBlocking));
    if (!c->edge_stream2) {
        int prio_lo = 0, prio_hi = 0;
        cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi);
        CK(cudaStreamCreateWithPriority(&c->edge_stream2, cudaStreamNonBlocking, prio_hi));
    }
    while ((int)c->band_up.size() < NB) {
        cudaEvent_t e1, e2, e3;
        CK(cudaEventCreateWithFlags(&e1, cudaEventDisableTiming)); c->band_up.push_back(e1);
        CK(cudaEventCreateWithFlags(&e2, cudaEventDisableTiming)); c->band_done.push_back(e2);
        CK(cudaEventCreateWithFlags(&e3, cudaEventDisableTiming)); c->band_edge.push_back(e3);
    }
    if (int rc = quiesce(c)) return rc;
    CK(cudaStreamSynchronize(c->copy_stream));
    if (zero) {
        for (int g = 0; g < c->ngen; ++g)
            for (int f = 0; f < 3; ++f)
                CK(cudaMemsetAsync(c->gen[g][f], 0, (size_t)c->lrows * c->pitch * c->esize, c->stream));
        c->cur = 0;
    }
    CK(cudaEventRecord(c->ev_t0, c->stream));
    const int G = (int)groups.size();
    const int gen0 = c->cur, gen1 = (c->cur + 1) % c->ngen;
    std::vector<int64_t> probe_base(G);
    for (int g = 0, acc = 0; g < G; ++g) { probe_base[g] = c->probe_count + acc; acc += groups[g].K; }
    auto upload = [&](int b) -> int {
        const int64_t r0 = (int64_t)b * H, r1 = std::min<int64_t>(r0 + H, NR);
        if (r1 > r0) {
            CK(cudaMemcpy2DAsync(plane<T>(c->epc, c, r0), c->pitch * 8, eps + (r0 - c->host_row0) * host_ld, host_ld * 8, NC * 8,
                                 r1 - r0, cudaMemcpyHostToDevice, c->copy_stream));
            CK(cudaMemcpy2DAsync(plane<T>(c->muc, c, r0), c->pitch * 8, mu + (r0 - c->host_row0) * host_ld, host_ld * 8, NC * 8,
                                 r1 - r0, cudaMemcpyHostToDevice, c->copy_stream));
        }
        CK(cudaEventRecord(c->band_up[b], c->copy_stream));
        return 0;
    };
    // Rows of h that are final once the LAST group has run on bands 0 .. b: a row r < (b+1) H belongs to a tile that starts
    // at or above it, i.e. in a band <= b.
    auto rows_done = [&](int b) -> int64_t { return b + 1 == NB ? NR : std::min<int64_t>((int64_t)(b + 1) * H, NR); };
    int64_t copied = 0;                                          // rows of h already on their way to the host
    auto download = [&](int b) -> int {
        if (!h_out) return 0;
        const int64_t upto = rows_done(b);
        if (upto <= copied) return 0;
        CK(cudaEventRecord(c->band_done[b], c->stream));         // the band's plain tiles of the last group ...
        CK(cudaStreamWaitEvent(c->d2h_stream, c->band_done[b], 0));
        CK(cudaStreamWaitEvent(c->d2h_stream, c->band_edge[b], 0));   // ... and its edge tiles
        const int final_gen = (G % 2 == 0) ? gen0 : gen1;
        CK(cudaMemcpy2DAsync(h_out + (copied - c->host_row0) * out_ld, out_ld * 8, plane<T>(c->gen[final_gen][0], c, copied),
                             c->pitch * 8, NC * 8, upto - copied, cudaMemcpyDeviceToHost, c->d2h_stream));
        copied = upto;
        return 0;
    };
    // diagonal sweep over groups [g0, g1): band + (group - g0) = d
    auto sweep = [&](int g0, int g1, bool up, bool down) -> int {
        const int ng = g1 - g0;
        for (int d = 0; d <= NB + ng; ++d) {
            if (up && d < NB) if (int rc = upload(d)) return rc;
            for (int k = 0; k < ng; ++k) {
                const int b = d - 1 - k, g = g0 + k;
                if (b < 0 || b >= NB) continue;
                if (up && k == 0) CK(cudaStreamWaitEvent(c->stream, c->band_up[std::min(b + 1, NB - 1)], 0));
                const int A = (g % 2 == 0) ? gen0 : gen1, B = (g % 2 == 0) ? gen1 : gen0;
                if (int rc = band_group<T>(c, groups[g].n, groups[g].K, A, B, b, NB, probe_base[g])) return rc;
                if (down && g == G - 1) if (int rc = download(b)) return rc;
            }
        }
        return 0;
    };
    auto join = [&]() -> int {                                   // the launch stream behind every edge tile queued so far
        for (int b = 0; b < NB; ++b) CK(cudaStreamWaitEvent(c->stream, c->band_edge[b], 0));
        return 0;
    };
    // A band unit is its own launch and fills the GPU only 1-2 times over, so swept groups run at about 0.4 of the speed of
    // whole-grid groups (B200: 16 ms instead of 6.2 ms per 8-call group on the C5 slab).  That is free while the copies
    // are the longer pole -- up to about 14 groups of 8 calls against the 0.23 s the three planes take over PCIe -- and a
    // loss beyond: long runs sweep only their first and last `wing` groups (200 calls: 295 vs 261 Gcell/s end to end).
    int wing = std::min(NB - 1, 4), sweep_all_upto = 14;
    if (const char* wg = std::getenv("FDTD_STREAM_WING")) { const int v = std::atoi(wg); if (v >= 1 && v <= 64) { wing = v; sweep_all_upto = 2 * v; } }
    int rc = 0;
    if (G <= std::max(sweep_all_upto, 2 * wing)) rc = sweep(0, G, true, true);
    else {
        rc = sweep(0, wing, true, false);
        if (!rc) rc = join();
        // every coefficient row has landed by now (the sweep's groups waited for their bands): the scan of the mu_coef
        // plane can run here, so that the whole-grid groups and the final sweep take the kernel without that stream
        if (!rc) rc = scan_mu_uniform(c);
        for (int g = wing; g < G - wing && !rc; ++g) {
            c->cur = (g % 2 == 0) ? gen0 : gen1;
            const int64_t pc = c->probe_count;
            c->probe_count = probe_base[g];
            rc = group_edge<T>(c, groups[g].n, groups[g].K);
            c->probe_count = pc;
            c->stats.steps -= groups[g].K;                       // accounted once, below
            c->stats.cell_updates -= (int64_t)groups[g].K * c->owned * c->cfg.cols;
        }
        if (!rc) rc = sweep(G - wing, G, false, true);
    }
    if (!rc) rc = join();
    if (rc) return rc;
    c->cur = (G % 2 == 0) ? gen0 : gen1;
    if (c->n_probe) c->probe_count += n_steps;
    c->stats.steps += n_steps;
    c->stats.cell_updates += n_steps * c->owned * c->cfg.cols;
    CK(cudaEventRecord(c->ev_t1, c->stream));
    if (int q = quiesce(c)) return q;
    CK(cudaStreamSynchronize(c->edge_stream2));
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(c->d2h_stream));
    float ms = 0.f;
    cudaEventElapsedTime(&ms, c->ev_t0, c->ev_t1);
    c->stats.last_run_ms = ms;
    return 0;
}

extern "C" {

int fdtd_run_streamed(fdtd_handle c, const double* eps_coef, const double* mu_coef, int64_t host_ld, int32_t zero_fields,
                      int64_t first_step, int64_t n_steps, double* h_out, int64_t out_ld) {
    NEED_BOUND();
    if (!eps_coef || !mu_coef || host_ld < c->cfg.cols || n_steps < 0 || (h_out && out_ld < c->cfg.cols))
        return fail(c, FDTD_ERR_ARG, "bad argument");
    // The pipeline serves the case it was built for: float64, the whole grid on this GPU, temporal blocking with stepK_edge.
    // Everything else takes the same four steps one after the other.
    const bool whole = c->cfg.row_begin == 0 && c->cfg.row_end == c->cfg.rows;
    if (c->cfg.dtype == FDTD_F64 && whole && !c->comm && c->temporal >= 2 && c->cfg.kernel == FDTD_KERNEL_STREAM &&
        !resident_eligible(c) && n_steps >= 2) {
        // the coefficient rows arrive while the first groups already run: nothing is known about the plane until the end
        // (long runs scan once all rows have landed, behind their first groups; short ones at the end, for the next run)
        c->mu_uniform = false;
        c->mu_scanned = false;
        const int rc = run_streamed<double>(c, eps_coef, mu_coef, host_ld, zero_fields != 0, first_step, n_steps, h_out, out_ld);
        if (rc != NOT_STREAMED) return rc ? rc : (c->mu_scanned ? 0 : scan_mu_uniform(c));
    }
    if (const char* strict = std::getenv("FDTD_STREAMED_STRICT"))            // test hook: the caller expected the pipeline
        if (std::atoi(strict)) return fail(c, FDTD_ERR_STATE, "fdtd_run_streamed: this run does not fit the pipeline");
    if (int rc = fdtd_upload_coeffs(c, eps_coef, mu_coef, host_ld)) return rc;
    if (zero_fields) if (int rc = fdtd_zero_fields(c)) return rc;
    if (int rc = fdtd_run(c, first_step, n_steps, 0, nullptr, 0, first_step + n_steps, nullptr)) return rc;
    if (h_out) {
        if (int rc = quiesce(c)) return rc;
        int rc;
        if (c->split) rc = split_d2h_field(c, c->gen[c->cur][0], h_out, out_ld, c->cfg.row_begin, c->owned, c->cfg.cols);
        else if (c->cfg.dtype == FDTD_F64) rc = d2h_plane<double>(c, c->gen[c->cur][0], h_out, out_ld, c->cfg.row_begin, c->owned, c->cfg.cols);
        else rc = d2h_plane<float>(c, c->gen[c->cur][0], h_out, out_ld, c->cfg.row_begin, c->owned, c->cfg.cols);
        if (rc) return rc;
        CK(cudaStreamSynchronize(c->stream));
    }
    return 0;
}

// ---- realtime view (the loop of solver.py:309-329 with batches of calls per frame) --------------------
int fdtd_view_begin(fdtd_handle c, int32_t row_stride, int32_t col_stride) {
    NEED_BOUND();
    Range nvtx("fdtd:view");
    if (row_stride < 1 || col_stride < 1) return fail(c, FDTD_ERR_ARG, "view strides must be >= 1");
    fdtd_ctx::View& v = c->view[c->view_head];
    if (v.pending) return fail(c, FDTD_ERR_STATE, "two views are in flight: fetch one with fdtd_view_end first");
    // rows of the GLOBAL picture (0, row_stride, 2 row_stride, ...) that this slab owns
    const int64_t first = (c->cfg.row_begin + row_stride - 1) / row_stride * row_stride;
    const int64_t rows = first < c->cfg.row_end ? (c->cfg.row_end - 1 - first) / row_stride + 1 : 0;
    const int64_t cols = (c->cfg.cols - 1) / col_stride + 1;
    const size_t bytes = (size_t)std::max<int64_t>(rows, 1) * cols * 8;
    if (bytes > v.cap) {
        if (v.dev) cudaFree(v.dev);
        if (v.pin) cudaFreeHost(v.pin);
        v.dev = v.pin = nullptr; v.cap = 0;
        CK(cudaMalloc(&v.dev, bytes));
        CK(cudaHostAlloc(&v.pin, bytes, cudaHostAllocDefault));
        v.cap = bytes;
    }
    if (!v.done) CK(cudaEventCreateWithFlags(&v.done, cudaEventDisableTiming));
    if (rows > 0) {
        dim3 block(64, 4), grid((unsigned)((cols + 63) / 64), (unsigned)std::min<int64_t>((rows + 3) / 4, 65535));
        if (c->split) {
            fdtd::PatchList none; none.n = 0;
            CK(fdtd::launch_split_to_f64(grid, block, c->stream, v.dev, cols, split_hi(c, c->gen[c->cur][0], first),
                                         split_lo(c, c->gen[c->cur][0], first), c->pitch, (int)rows, (int)cols, row_stride,
                                         col_stride, 0, none));
        } else if (c->cfg.dtype == FDTD_F64)
            CK(fdtd::launch_plane_view<double>(grid, block, c->stream, v.dev, cols, plane<double>(c->gen[c->cur][0], c, first),
                                               c->pitch, (int)rows, (int)cols, row_stride, col_stride));
        else
            CK(fdtd::launch_plane_view<float>(grid, block, c->stream, v.dev, cols, plane<float>(c->gen[c->cur][0], c, first),
                                              c->pitch, (int)rows, (int)cols, row_stride, col_stride));
        c->stats.kernel_launches++;
        // the copy leaves on the side stream: the next batch of calls, queued behind the pack kernel, runs beside it
        CK(cudaEventRecord(v.done, c->stream));
        CK(cudaStreamWaitEvent(c->copy_stream, v.done, 0));
        CK(cudaMemcpyAsync(v.pin, v.dev, (size_t)rows * cols * 8, cudaMemcpyDeviceToHost, c->copy_stream));
    }
    CK(cudaEventRecord(v.done, c->copy_stream));
    // the pack kernel of the view after next must not overwrite v.dev before this copy has read it: that kernel is
    // queued on c->stream only after fdtd_view_end synchronised on v.done, so no extra dependency is needed
    v.rows = rows; v.cols = cols; v.pending = true;
    c->view_head ^= 1;
    return 0;
}

int fdtd_view_end(fdtd_handle c, double* host_out, int64_t capacity, int64_t* rows, int64_t* cols) {
    NEED_BOUND();
    fdtd_ctx::View& v = c->view[c->view_tail];
    if (!v.pending) return fail(c, FDTD_ERR_STATE, "no view in flight");
    if (!host_out || capacity < v.rows * v.cols) return fail(c, FDTD_ERR_ARG, "view buffer too small");
    CK(cudaEventSynchronize(v.done));
    std::memcpy(host_out, v.pin, (size_t)(v.rows * v.cols) * 8);
    if (rows) *rows = v.rows;
    if (cols) *cols = v.cols;
    v.pending = false;
    c->view_tail ^= 1;
    return 0;
}

int fdtd_set_tuning(fdtd_handle c, int32_t tile_rows, int32_t vec, int32_t block_warps) {
    if (!c) return FDTD_ERR_ARG;
    const int va = c->cfg.dtype != FDTD_F32 ? 2 : 4;
    if (vec && vec != va && vec != 2 * va) return fail(c, FDTD_ERR_ARG, "unsupported vec (2/4 for f64, 4/8 for f32)");
    if (block_warps < 0 || block_warps > 8 || tile_rows < 0) return fail(c, FDTD_ERR_ARG, "bad tuning");
    if (tile_rows) { c->cfg.tile_rows = tile_rows; c->auto_tile_rows = false; }
    if (vec) c->cfg.vec = vec;
    if (block_warps) { c->cfg.block_warps = block_warps; c->auto_warps = false; }
    for (FrameMaps& m : c->fm) m.key.clear();
    return 0;
}

int fdtd_get_tuning(fdtd_handle c, int32_t out[8]) {
    if (!c || !out) return FDTD_ERR_ARG;
    const int v[8] = {c->cfg.tile_rows, c->cfg.vec, c->cfg.block_warps, c->temporal, c->frame_mode, c->kstep_variant,
                      c->halo_depth, c->ngen};
    for (int k = 0; k < 8; ++k) out[k] = v[k];
    return 0;
}

int fdtd_kernel_timing(fdtd_handle c, int32_t enable) {
    if (!c) return FDTD_ERR_ARG;
    CK(cudaSetDevice(c->cfg.device));
    if (enable && c->kt_ev.empty()) {
        c->kt_ev.resize(2 * 4096);
        for (cudaEvent_t& e : c->kt_ev) CK(cudaEventCreate(&e));
    }
    c->ktime_on = enable != 0;
    c->kt_n = 0;
    c->kt_cell_steps = 0;
    return 0;
}

int fdtd_kernel_timing_fetch(fdtd_handle c, int32_t* n_launches, double* total_ms, double* min_ms, double* max_ms,
                             int64_t* cell_updates) {
    if (!c) return FDTD_ERR_ARG;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->edge_stream));
    double tot = 0, lo = 0, hi = 0;
    for (int k = 0; k < c->kt_n; ++k) {
        float ms = 0.f;
        CK(cudaEventElapsedTime(&ms, c->kt_ev[2 * k], c->kt_ev[2 * k + 1]));
        tot += ms;
        lo = k == 0 ? ms : std::min<double>(lo, ms);
        hi = std::max<double>(hi, ms);
    }
    if (n_launches) *n_launches = c->kt_n;
    if (total_ms) *total_ms = tot;
    if (min_ms) *min_ms = lo;
    if (max_ms) *max_ms = hi;
    if (cell_updates) *cell_updates = c->kt_cell_steps;
    return 0;
}

int fdtd_sync(fdtd_handle c) {
    if (!c) return FDTD_ERR_ARG;
    CK(cudaSetDevice(c->cfg.device));
    CK(cudaStreamSynchronize(c->stream));
    CK(cudaStreamSynchronize(c->copy_stream));
    CK(cudaStreamSynchronize(c->comm_stream));
    CK(cudaStreamSynchronize(c->edge_stream));
    return 0;
}

int fdtd_get_stats(fdtd_handle c, fdtd_stats* out) {
    if (!c || !out) return FDTD_ERR_ARG;
    *out = c->stats;
    return 0;
}

int fdtd_current_generation(fdtd_handle c) { return c ? c->cur : FDTD_ERR_ARG; }

int fdtd_comm_unique_id(void* out128) {
    fdtd_ctx* c = nullptr;
    if (!out128) return fail(c, FDTD_ERR_ARG, "null id buffer");
    std::string err;
    if (!g_nccl.load(err)) return fail(c, FDTD_ERR_COMM, err);
    NcclId id;
    NCK(g_nccl.GetUniqueId(&id));
    std::memcpy(out128, id.internal, 128);
    return 0;
}

int fdtd_comm_init(fdtd_handle c, const void* unique_id128, int32_t rank, int32_t nranks) {
    if (!c || !unique_id128) return FDTD_ERR_ARG;
    if (nranks < 1 || rank < 0 || rank >= nranks) return fail(c, FDTD_ERR_ARG, "bad rank/nranks");
    if (c->split) return fail(c, FDTD_ERR_ARG, "split storage (FDTD_F32X) runs on one GPU");
    std::string err;
    if (!g_nccl.load(err)) return fail(c, FDTD_ERR_COMM, err);
    CK(cudaSetDevice(c->cfg.device));
    NcclId id;
    std::memcpy(id.internal, unique_id128, 128);
    NCK(g_nccl.CommInitRank(&c->comm, nranks, id, rank));
    c->rank = rank; c->nranks = nranks;
    return 0;
}

int fdtd_halo_exchange(fdtd_handle c) {
    NEED_BOUND();
    if (!c->comm) return 0;
    CK(cudaStreamSynchronize(c->stream));
    if (int rc = exchange(c, c->cur, c->comm_stream)) return rc;
    CK(cudaStreamSynchronize(c->comm_stream));
    return 0;
}

int fdtd_halo_rows(fdtd_handle c, void* send_down[3], void* recv_up[3], void** send_up, void** recv_down,
                   int64_t counts[3]) {
    NEED_BOUND();
    const int g = c->cur;
    auto row = [&](int f, int64_t grow) { return (void*)((char*)c->gen[g][f] + (size_t)((grow - c->row_off) * c->pitch) * c->esize); };
    for (int f = 0; f < 3; ++f) {
        if (send_down) send_down[f] = row(f, c->cfg.row_end - 1);
        if (recv_up) recv_up[f] = c->halo_top ? row(f, c->cfg.row_begin - 1) : nullptr;
    }
    if (send_up) *send_up = row(1, c->cfg.row_begin);
    if (recv_down) *recv_down = row(1, c->cfg.row_end);
    if (counts) { counts[0] = c->cfg.cols; counts[1] = c->cfg.cols; counts[2] = c->cfg.cols + 1; }
    return 0;
}

int fdtd_halo_blocks(fdtd_handle c, void* send_down[3], void* recv_down[3], void* send_up[3], void* recv_up[3],
                     int64_t* count) {
    NEED_BOUND();
    if (c->halo_depth <= 1) return fail(c, FDTD_ERR_STATE, "fdtd_halo_blocks needs deep halos (fdtd_set_halo_depth > 1)");
    const int g = c->cur;
    const int64_t D = c->halo_depth;
    auto row = [&](int f, int64_t grow) { return (void*)((char*)c->gen[g][f] + (size_t)((grow - c->row_off) * c->pitch) * c->esize); };
    const bool up = c->cfg.row_begin > 0, down = c->cfg.row_end < c->cfg.rows;
    for (int f = 0; f < 3; ++f) {
        if (send_down) send_down[f] = down ? row(f, c->cfg.row_end - D) : nullptr;
        if (recv_down) recv_down[f] = down ? row(f, c->cfg.row_end) : nullptr;
        if (send_up) send_up[f] = up ? row(f, c->cfg.row_begin) : nullptr;
        if (recv_up) recv_up[f] = up ? row(f, c->cfg.row_begin - D) : nullptr;
    }
    if (count) *count = D * c->pitch;
    return 0;
}

}  // extern "C"
