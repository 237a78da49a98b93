/* fdtd2d.h -- C ABI of the B200-native 2D Yee time-stepper (libfdtd2d.so).
 *
 * The reference (jmanchuck/FDTD-EM-Solver) has no FFI layer: its hot path is the
 * Python method Solver.update (solver.py:169-257) driven by the loops in
 * Solver.solve (solver.py:288-292, :331-337).  This header is the boundary a
 * maintainer binds (ctypes/cffi) to replace exactly that
 * path; each entry point cites the reference lines it stands in for.
 *
 * Conventions: every function returns 0 on success and a negative fdtd_status
 * on failure; fdtd_last_error() gives the message.  No C++ exceptions, no
 * exit(), no torch types.  Device buffers are owned by the caller (raw device
 * pointers, e.g. torch tensors' data_ptr()); the library owns only its small
 * parameter tables, streams/events and snapshot staging.
 *
 * Geometry.  Global grid: h (rows, cols), ex (rows+1, cols), ey (rows, cols+1)
 * as in solver.py:101-103, but rows != cols is allowed.  A context owns the
 * y-slab of global rows [row_begin,row_end).  Every device array of a context
 * has the same shape (fdtd_local_rows() x pitch elements, C order):
 *   local row l  <->  global row  row_begin - halo_top + l,
 *   halo_top = (row_begin > 0) ? 1 : 0       (old h/ex/ey/mu_coef of the row above),
 *   the last local row holds ex[row_end] (halo from the slab below, or the
 *   real last ex row when row_end == rows).
 * pitch >= cols+1 and a multiple of 32 elements (fdtd_default_pitch()).
 */
#ifndef FDTD2D_H
#define FDTD2D_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fdtd_ctx* fdtd_handle;

enum fdtd_status {
    FDTD_OK = 0,
    FDTD_ERR_ARG = -1,      /* bad argument / unsupported geometry          */
    FDTD_ERR_CUDA = -2,     /* a CUDA runtime call failed                   */
    FDTD_ERR_STATE = -3,    /* call order (buffers not bound, ...)          */
    FDTD_ERR_RANGE = -4,    /* step outside the uploaded source tables      */
    FDTD_ERR_COMM = -5      /* NCCL halo transport failed / unavailable     */
};

/* FDTD_F64: the parity mode (bit-identical to the reference).  FDTD_F32: everything in float32, 32 bytes per cell-update, the
 * throughput mode; its error grows with the length of the run (field-STORAGE rounding) and leaves the 1e-5 gate on long
 * runs.  FDTD_F32X ("float32x"): the fields are stored as an fp32 "hi" plane plus a bf16 "lo" plane (6 bytes per value, 32
 * significant bits: each field buffer holds local_rows x pitch floats followed by local_rows x pitch uint16), the
 * coefficient planes, the pulse tables and ALL arithmetic are float64 in the reference's operation order; 52 bytes per
 * cell-update (3 fields x 6 B x 2 + 2 coefficients x 8 B) instead of 64, 1e-5 relative L2 held on every snapshot of the
 * reference's scripts.  One GPU, one kernel family (stepK_edge with every feature, up to 8 calls per launch, the bits of
 * single calls); no TF/SF "right" line at column 0. */
enum fdtd_dtype { FDTD_F64 = 0, FDTD_F32 = 1, FDTD_F32X = 2 };

/* Pulse.get_direction(): None / "right" / "left" / "up" / "down" (solver.py:484, :176-217). */
enum fdtd_source_kind { FDTD_SRC_POINT = 0, FDTD_SRC_RIGHT = 1, FDTD_SRC_LEFT = 2, FDTD_SRC_UP = 3, FDTD_SRC_DOWN = 4 };

enum fdtd_kernel {
    FDTD_KERNEL_EXACT = 0,   /* one thread per cell, closed-form restatement of solver.py:169-242 (in-repo GPU reference) */
    FDTD_KERNEL_STREAM = 1   /* fused H+E streaming kernel, the production path                                         */
};

typedef struct fdtd_config {
    int64_t rows, cols;            /* solver.py:98 `size` (both), rectangular allowed               */
    int64_t row_begin, row_end;    /* owned y-slab; 0,rows on one GPU                               */
    int64_t pitch;                 /* elements per row of every device array; 0 = default           */
    int32_t dtype;                 /* fdtd_dtype                                                    */
    int32_t device;                /* CUDA device ordinal                                           */
    int32_t reflect[4];            /* solver.py:54 boundaries [up,down,left,right], 1=reflect       */
    int32_t kernel;                /* fdtd_kernel                                                   */
    int32_t tile_rows;             /* rows marched per warp tile (0 = default)                      */
    int32_t vec;                   /* elements per thread per row (0 = default; 2/4 f64, 4/8 f32)   */
    int32_t block_warps;           /* warps per CTA (0 = default)                                   */
    double courant;                /* (C*dt/ds) evaluated by the host in Python floats, solver.py:221 */
    double one_minus_courant;      /* (1 - (C*dt/ds)), same                                         */
} fdtd_config;

typedef struct fdtd_source {
    int32_t kind;                  /* fdtd_source_kind                                              */
    int32_t reserved;
    int64_t row, col;              /* Pulse.get_row()/get_col(), solver.py:635-639 (global indices) */
} fdtd_source;

/* One MaterialArray painting call, already converted to grid indices (solver.py:368-466), for fdtd_paint_coeffs.
 *   FDTD_SHAPE_RECT    a = {r0, r1, c0, c1}: rows [r0,r1), cols [c0,c1) after Python slice clipping of
 *                      eps_arr[ul0:lr0, ul1:lr1+1] (solver.py:374)
 *   FDTD_SHAPE_QUARTER a = {ci, cj, radius, thickness}: radius^2 > x^2+y^2 > (radius-thickness)^2, x<0, y>0 (solver.py:426-429)
 *   FDTD_SHAPE_CONVEX  a = {ci, cj, radius, displacer, size}: the two caps of solver.py:393-409 incl. the wrap of
 *                      negative indices and the skipping of indices >= size                                        */
enum fdtd_shape_kind { FDTD_SHAPE_RECT = 0, FDTD_SHAPE_QUARTER = 1, FDTD_SHAPE_CONVEX = 2 };
typedef struct fdtd_shape {
    int32_t kind;
    int32_t reserved;
    int64_t a[6];
    double eps;                    /* absolute: epsilon_rel * EPSILON, evaluated by the host */
    double mu;                     /* absolute: mu_rel * MU                                  */
} fdtd_shape;

typedef struct fdtd_stats {
    int64_t steps;                 /* leapfrog steps executed since create                          */
    int64_t cell_updates;          /* owned rows * cols * steps                                     */
    int64_t kernel_launches;       /* kernels of this library launched since create                 */
    int64_t halo_bytes_sent;       /* NVLink halo traffic, this rank                                */
    double  last_run_ms;           /* device time of the last fdtd_run (CUDA events)                */
} fdtd_stats;

int64_t fdtd_default_pitch(int64_t cols, int32_t dtype);
int64_t fdtd_local_rows(const fdtd_config* cfg);

/* Solver.update_constants() field allocation + solve() preamble (solver.py:93-105, :270-286). */
int fdtd_create(const fdtd_config* cfg, fdtd_handle* out);
int fdtd_destroy(fdtd_handle h);
const char* fdtd_last_error(fdtd_handle h);          /* h may be NULL: last error of a failed fdtd_create */

/* Device arrays (caller-owned, fdtd_local_rows() x pitch each, zero-filled by the caller or by
 * fdtd_zero_fields): two generations of h/ex/ey (ping-pong) and the two coefficient planes. */
int fdtd_bind_buffers(fdtd_handle h, void* h0, void* h1, void* ex0, void* ex1, void* ey0, void* ey1,
                      void* eps_coef, void* mu_coef);

/* Temporal blocking (BASELINE.json north_star (1)): fdtd_run advances K leapfrog steps per HBM pass.  Tiles whose K-step
 * dependency cone touches no wall, source, reflector, probe or slab edge go through the plain K-step kernel (stepK_lean);
 * all other tiles go through stepK_edge, the same register pipeline with those features folded in as tile predicates
 * (frame mode 3, the default) -- two launches per group that read generation A and write generation B, no scratch.
 * Scratch generations of h/ex/ey (fdtd_bind_scratch: once for 2 steps per pass, twice for 3..8) are needed only by the
 * fallback, frame mode 0: K masked single-step launches through the scratch generations for the non-plain tiles, used
 * when stepK_edge cannot take the scene (a TF/SF "right" line at column 0, more than 16 pulses) or a y-slab stores
 * fewer than K+1 halo rows (fdtd_set_halo_depth); without scratch such runs step call by call.  Results are
 * bit-identical to single stepping in every mode.  Groups never straddle a snapshot. */
int fdtd_bind_scratch(fdtd_handle h, void* h2, void* ex2, void* ey2);
int fdtd_set_temporal_blocking(fdtd_handle h, int32_t steps_per_pass);

/* Resident mode for grids that live in L2 (the reference's own scenario scripts, N = 139..442; solver.py:288-292 and
 * :331-337 drive thousands of update() calls on them): with enable != 0, fdtd_run advances ALL calls up to the next
 * snapshot in ONE cooperative launch of step_resident (fields ping-pong in L2, per-call pulse records read from a
 * device table, grid-wide barrier between calls) instead of one launch per call, and fdtd_step launches the same
 * kernel for its one call.  Used when the context owns the
 * whole grid, (rows+1)*(cols+1) <= max_cells (0 keeps the default 512*512), no halo transport is active and no TF/SF
 * "right" line sits at column 0; otherwise fdtd_run / fdtd_step silently keep the per-call path.  Results are bit-identical
 * either way.  fdtd_resident_active reports whether fdtd_run would use it for the current scene (1/0). */
int fdtd_set_resident(fdtd_handle h, int32_t enable, int64_t max_cells);
int fdtd_resident_active(fdtd_handle h);
/* Which K-step kernel fdtd_run launches for the plain tiles of a temporal-blocking group: 1 = stepK_lean (default): the
 * stage pipeline with the row loop unrolled by two and running pointers instead of per-row address arithmetic (about a
 * third fewer issue slots per row); 0 = stepK_stream, the first version of the same pipeline.  Same operations in the
 * same order: bit-identical. */
int fdtd_set_kstep_variant(fdtd_handle h, int32_t variant);

/* How the tiles the plain K-step kernel leaves out (the FRAME: wall bands and columns, pulse lines, sources, reflectors,
 * probes) advance in a temporal-blocking group:
 *   3  ONE launch of stepK_edge (csrc/edge.cuh) next to the plain kernel (default);
 *   0  K masked single-step launches through the scratch generations (needs fdtd_bind_scratch).
 * Mode 3 does not apply (the group silently uses mode 0, or single steps without scratch) when a TF/SF "right" line
 * sits at column 0 or more than 16 pulses exist, and on a y-slab it needs K+1 halo rows (fdtd_set_halo_depth).  Results
 * are bit-identical in both modes.  (Round 2 also measured a shared-memory variant, "frame_window", as modes 1/2: bit-exact
 * but slower than stepK_edge on the B200 -- 512 vs 552 Gcell/s, profiles/r2_sweep_b_summary.txt -- and removed.) */
int fdtd_set_frame_mode(fdtd_handle h, int32_t mode);

/* The frame plan of temporal blocking as pure host logic (no CUDA call, works without a GPU; used by the tests).
 * cfg: rows, cols, row_begin, row_end, dtype, tile_rows (> 0) and block_warps (> 0) are read.  boxes: n_boxes x
 * {r0, r1, c0, c1} INCLUSIVE cell boxes that the K-step kernel must stay away from (sources, reflectors, probes).
 * plan receives the geometry; skip (k_tiles_y x k_tiles_x bytes, 1 = tile left to the frame passes) and masks
 * (steps_per_pass x frame_tiles_y x frame_tiles_x bytes, masks[d] = frame dilated by d rings of tiles) may be NULL. */
typedef struct fdtd_frame_plan {
    int32_t k_tile_rows, k_tile_cols, k_load_cols, k_margin_cols;   /* stepK_stream tile: rows, stored / loaded columns, halo */
    int32_t k_tiles_y, k_tiles_x;
    int64_t k_row_lo, k_row_hi;                                     /* rows owned by K-step tiles: [k_row_lo, k_row_hi)    */
    int32_t frame_tile_rows, frame_tile_cols, frame_tiles_y, frame_tiles_x;
} fdtd_frame_plan;
int fdtd_plan_frame(const fdtd_config* cfg, int32_t steps_per_pass, int32_t n_boxes, const int64_t* boxes,
                    fdtd_frame_plan* plan, uint8_t* skip, uint8_t* masks);

/* Host arrays handed to fdtd_upload_coeffs / fdtd_set_fields / fdtd_get_fields normally start at global
 * row 0.  A rank that keeps only its slab on the host declares the global row its host arrays start at. */
int fdtd_set_host_row_origin(fdtd_handle h, int64_t first_global_row);

/* eps_coefficient_arr / mu_coefficient_arr (solver.py:278-279), GLOBAL host arrays (rows x cols,
 * leading dimension host_ld doubles); the context takes the rows it needs (incl. halo_top). */
int fdtd_upload_coeffs(fdtd_handle h, const double* eps_coef, const double* mu_coef, int64_t host_ld);

/* C5 synthetic coefficients generated on the device: eps_r = 1 + 3*u(i,j),
 * u = (splitmix64(seed + i*cols + j) >> 11) * 2^-53, mu_r = 1; same bits as oracle.restated.synthetic_coefficients. */
int fdtd_fill_synthetic_coeffs(fdtd_handle h, uint64_t seed, double dt_over_ds, double epsilon0, double mu0);

/* MaterialArray painting + coefficient build on the device, for grids whose eps/mu arrays do
 * not fit in host RAM: the shapes are applied in list order over the background, then
 * eps_coef = dt_over_ds * (1 / eps), mu_coef = dt_over_ds * (1 / mu) exactly as solver.py:278-279 rounds them. */
int fdtd_paint_coeffs(fdtd_handle h, int32_t n, const fdtd_shape* shapes, double eps_background, double mu_background,
                      double dt_over_ds);

/* One property of the coefficient planes is remembered by the library: a UNIFORM mu_coef plane (mu_r = 1 everywhere, as in
 * every script of the reference and in the C5 recipe) lets the plain K-step kernel read the value from its parameters
 * instead of streaming the plane (7 words per cell and pass instead of 8; same bits).  fdtd_upload_coeffs,
 * fdtd_fill_synthetic_coeffs, fdtd_paint_coeffs and fdtd_run_streamed establish it with one device pass over the plane;
 * a caller that writes the caller-owned coefficient tensors directly calls fdtd_rescan_coeffs afterwards.
 * fdtd_uniform_mu reports it (1/0).  FDTD_UNIFORM_MU=0 in the environment keeps the general kernel. */
int fdtd_rescan_coeffs(fdtd_handle h);
int fdtd_uniform_mu(fdtd_handle h);

/* self.h / self.ex / self.ey (solver.py:101-103).  GLOBAL host arrays, any pointer may be NULL
 * (set: zero-fill that field; get: skip it).  set also loads the halo rows; get writes owned rows only. */
int fdtd_set_fields(fdtd_handle h, const double* hh, const double* ex, const double* ey);
int fdtd_get_fields(fdtd_handle h, double* hh, double* ex, double* ey);
int fdtd_zero_fields(fdtd_handle h);                 /* Solver.reset(), solver.py:347-349 */

/* Pulses (solver.py:176-217).  Tables are host arrays evaluated by the caller with the reference's
 * own expressions: mag[p*table_len+n] = pulse.magnitude(n); mag_z = mag*sqrt(MU/EPSILON);
 * active[p*table_len+n] = (start < time_vector[n] < end); last_mag[n] = magnitude of the last active
 * pulse in list order (the stale `magnitude` local of solver.py:178/:196-200). */
int fdtd_set_sources(fdtd_handle h, int32_t n, const fdtd_source* src, int64_t table_len,
                     const double* mag, const double* mag_z, const uint8_t* active, const double* last_mag);

/* self.reflectors (solver.py:114-120, :164-167, :238-242): n x [r0,r1,c0,c1] with Python slice semantics. */
int fdtd_set_reflectors(fdtd_handle h, int32_t n, const int64_t* rects);

/* record_energy cells (solver.py:343-345): raw (h,ex,ey) triples captured after every step (:246-255). */
int fdtd_set_probes(fdtd_handle h, int32_t n, const int64_t* ij);
int fdtd_probe_fetch(fdtd_handle h, double* out, int64_t max_records, int64_t* n_records);
int fdtd_probe_clear(fdtd_handle h);

/* One Solver.update(time) call for call index `step` (= self.step on entry), asynchronous. */
int fdtd_step(fdtd_handle h, int64_t step);

/* The solve() loops (solver.py:288-292 and :331-337): steps [first_step, first_step+n_steps).
 * If snapshot_every > 0, after every call with (step+1) % snapshot_every == 0 the owned rows of h are
 * written as float64 to snapshots_host[k] (owned_rows x cols each, dense), k counting from 0, with the
 * reference's list-aliasing rule applied (solver.py:337 vs :180: point-source cells hold the NEXT call's
 * magnitude when that call exists, i.e. step+1 < total_calls, and the pulse is active then).
 * snapshots_host may be ordinary pageable memory: each snapshot goes device -> a pinned ring owned by the library
 * (async DMA on a side stream) -> snapshots_host (memcpy in stream order).  Returns after everything completed. */
int fdtd_run(fdtd_handle h, int64_t first_step, int64_t n_steps, int64_t snapshot_every,
             double* snapshots_host, int64_t max_snapshots, int64_t total_calls, int64_t* n_snapshots);

/* The realtime view (the reference's plot_and_show loop, solver.py:309-329, with batches of calls per frame).
 * fdtd_run_enqueue queues calls [first_step, first_step + n_steps) exactly as fdtd_run does (no snapshots) and returns
 * without waiting; queued runs execute back to back.  fdtd_view_begin queues, behind everything queued so far, a picture
 * of h -- global rows 0, row_stride, 2 row_stride, ... that this slab owns, columns 0, col_stride, ... -- as float64:
 * packed on the launch stream, copied to pinned memory on the side stream, so the copy overlaps whatever is queued next.
 * fdtd_view_end waits for the OLDEST picture in flight and copies it to host_out (capacity in doubles; *rows x *cols,
 * dense).  At most two pictures are in flight.  Typical frame: run_enqueue(batch k+1); view_begin(); view_end() -> the
 * picture of batch k is drawn while the device computes batch k+1.  No aliasing patch is applied: the realtime loop
 * draws the field update() returned, not the saved list (solver.py:323 vs :337). */
int fdtd_run_enqueue(fdtd_handle h, int64_t first_step, int64_t n_steps);
/* fdtd_upload_coeffs + (zero_fields: fdtd_zero_fields) + fdtd_run(first_step, n_steps) + fdtd_get_fields(h) as ONE pipeline
 * over row bands: band b's coefficient rows travel to the device while the calls already run on the bands above it, and a
 * band's rows of h travel back behind its last call while the bands below are still stepping -- PCIe busy in both
 * directions, the stepping hidden behind it (the whole-job path of Solver.solve(solve_only=True) followed by a read of
 * solver.h, solver.py:278-292).  h_out (may be NULL: no download) receives the owned rows of h, out_ld doubles per row,
 * addressed like fdtd_get_fields (see fdtd_set_host_row_origin).  Pinned host memory makes the copies asynchronous;
 * pageable memory works, without the overlap.  Pipelined for float64, the whole grid on one GPU, temporal blocking >= 2;
 * any other configuration runs the four steps one after the other.  Bit-identical to them either way. */
int fdtd_run_streamed(fdtd_handle h, const double* eps_coef, const double* mu_coef, int64_t host_ld, int32_t zero_fields,
                      int64_t first_step, int64_t n_steps, double* h_out, int64_t out_ld);

int fdtd_view_begin(fdtd_handle h, int32_t row_stride, int32_t col_stride);
int fdtd_view_end(fdtd_handle h, double* host_out, int64_t capacity, int64_t* rows, int64_t* cols);

/* Launch geometry of the streaming kernel (0 keeps the current value); results never depend on it. */
int fdtd_set_tuning(fdtd_handle h, int32_t tile_rows, int32_t vec, int32_t block_warps);
/* What the library runs with (the defaults it picked included): out = {tile_rows, vec, block_warps, steps_per_pass,
 * frame_mode, kstep_variant, halo_depth, field generations bound}. */
int fdtd_get_tuning(fdtd_handle h, int32_t out[8]);

/* Measurement hook: with enable != 0 every launch of the plain K-step kernel (the dominant kernel of fdtd_run with
 * temporal blocking) is bracketed by CUDA events on the stream it is launched on (at most 4096 launches are kept; the
 * counters restart at every call; only launches of the configured group size count).  fdtd_kernel_timing_fetch waits for
 * the streams and returns the number of launches timed, their total, shortest and longest duration in ms, and the
 * cell-updates (cells x steps) those launches advanced. */
int fdtd_kernel_timing(fdtd_handle h, int32_t enable);
int fdtd_kernel_timing_fetch(fdtd_handle h, int32_t* n_launches, double* total_ms, double* min_ms, double* max_ms,
                             int64_t* cell_updates);

int fdtd_sync(fdtd_handle h);
int fdtd_get_stats(fdtd_handle h, fdtd_stats* out);

/* Which generation (0/1) currently holds the fields -- for callers that read the torch tensors directly. */
int fdtd_current_generation(fdtd_handle h);

/* y-slab halo exchange over NCCL (no reference counterpart).  One context per rank.
 * fdtd_comm_unique_id fills a 128-byte ncclUniqueId on rank 0; broadcast it out-of-band
 * (torch.distributed) and pass it to fdtd_comm_init on every rank.  Afterwards fdtd_step/fdtd_run
 * exchange 3 rows down / 1 row up per step, overlapped with the interior update. */
int fdtd_comm_unique_id(void* out128);
int fdtd_comm_init(fdtd_handle h, const void* unique_id128, int32_t rank, int32_t nranks);
/* Exchange the halo rows of the CURRENT generation (after fdtd_set_fields on every rank). */
int fdtd_halo_exchange(fdtd_handle h);
/* Row pointers/counts of the current generation for an external transport (torch.distributed):
 * send_down[3] = owned last rows of h,ex,ey; recv_up[3] = halo_top rows; send_up = ex[row_begin];
 * recv_down = ex[row_end] slot.  counts in elements. */
int fdtd_halo_rows(fdtd_handle h, void* send_down[3], void* recv_up[3], void** send_up, void** recv_down,
                   int64_t counts[3]);

/* Deep halos (BASELINE.json north_star (3): "k-row halo exchange under temporal blocking").  With depth D >= 2 a slab
 * stores D rows of h/ex/ey (and of the coefficient planes) of EACH neighbouring slab:
 *   local row l  <->  global row  row_begin - (row_begin > 0 ? D : 0) + l,   fdtd_local_rows_halo() rows in all,
 * so a temporal-blocking group of K <= D-1 steps needs nothing from the neighbours while it runs: the K-step kernel
 * reaches the slab's first/last owned row, stepK_edge takes every other tile (frame mode 3), and the D-row blocks
 * are exchanged ONCE per group (one contiguous D x pitch block per array and direction) instead of one row per step and
 * masked pass.  Call fdtd_set_halo_depth right after fdtd_create, before the buffers are sized and bound; every slab
 * next to an interface must own >= D rows.  Results are bit-identical to depth 1. */
int64_t fdtd_local_rows_halo(const fdtd_config* cfg, int32_t halo_depth);
int fdtd_set_halo_depth(fdtd_handle h, int32_t halo_depth);
/* The tile plan of frame mode 3 as pure host logic (no CUDA call).  Every tile of the slab is a K-step tile: skip
 * (k_tiles_y x k_tiles_x bytes) = 0 for the tiles of the plain K-step kernel; every other tile is listed in edge_tiles
 * as {column tile tx, first row, end row, next-to-an-interface flag | tile class << 8} for stepK_edge (csrc/edge.cuh;
 * the class is the feature set of the instantiation that runs the tile: 1 grid rows 0 / NR-1 / NR or unstored rows,
 * 2 grid columns 0 / NC-1 / NC, 4 pulses, 8 reflectors and probes), which folds walls,
 * pulses, reflectors and probes into the same register pipeline -- incl. the K-row bands at the top / bottom wall that
 * the plain tile grid (rows k_row_lo .. k_row_hi) leaves out.  A y-slab needs halo_depth >= steps_per_pass + 1.
 * box_is_pulse (NULL, or one flag per box) tells a pulse's box (class bit 4) from a reflector / probe box (class bit 8);
 * NULL counts every box as both.  *n_edge_tiles receives the number of edge tiles; at most max_edge_tiles quadruples are
 * written to edge_tiles, in launch order (tiles of one class are consecutive within the near and the far half). */
int fdtd_plan_edge(const fdtd_config* cfg, int32_t steps_per_pass, int32_t halo_depth, int32_t n_boxes, const int64_t* boxes,
                   const int32_t* box_is_pulse, fdtd_frame_plan* plan, uint8_t* skip, int32_t* n_edge_tiles,
                   int32_t* edge_tiles, int32_t max_edge_tiles);
/* Block pointers of the current generation for an external transport: send_down = the last D owned rows, recv_down = the
 * D rows stored from the slab below, send_up = the first D owned rows, recv_up = the D rows stored from the slab above
 * (NULL where there is no neighbour); *count = elements per block (D x pitch). */
int fdtd_halo_blocks(fdtd_handle h, void* send_down[3], void* recv_down[3], void* send_up[3], void* recv_up[3],
                     int64_t* count);

#ifdef __cplusplus
}
#endif
#endif /* FDTD2D_H */
