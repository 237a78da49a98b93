"""Device engine: torch tensors as device buffers + the libfdtd2d C ABI.

This is the layer the drop-in `Solver` (solver.py) and bench.py sit on.  It
replaces the arrays and loops of the reference's Solver.solve / Solver.update
(/root/reference/solver.py:169-257, :270-337); all arithmetic happens in the
CUDA library.  There is deliberately no CPU implementation here.
"""
import ctypes as C
import math
import os

import numpy as np

from . import _lib as L

# solver.py:9-11, same expressions -> same float64 bits
C0 = 2.99 * (10 ** 8)
MU = 1.26 * (10 ** (-6))
EPSILON = 8.85 * (10 ** (-12))
IMPEDANCE = math.sqrt(MU / EPSILON)


# Resident mode (many update() calls per cooperative launch, for grids that live in L2): what Engine(resident=None)
# means when the caller also left the launch geometry and temporal blocking alone.  FDTD_RESIDENT=0/1 overrides it.
# On since its first B200 run (profiles/r1_resident_*): bit-identical, 7-9 us per call instead of 43-63 us.
RESIDENT_DEFAULT = True


def _resident_default():
    v = os.environ.get("FDTD_RESIDENT")
    return RESIDENT_DEFAULT if v is None or v == "" else v not in ("0", "false", "no")


def empty_hugepages(shape):
    """np.empty(shape, float64) whose pages the kernel may back with transparent huge pages (MADV_HUGEPAGE).  A snapshot
    stack is written exactly once, front to back, by the library's copy thread, so its cost is first-touch page faults:
    with THP in `madvise` mode (this image) huge pages make that 3.4x faster (1.07 -> 3.6 GB/s on the build host,
    measured with 814 KB snapshots).  Any failure is ignored: it is only a hint."""
    a = np.empty(shape, dtype=np.float64)
    if a.nbytes >= (8 << 20):
        try:
            page = 4096
            lo = (a.ctypes.data + page - 1) & ~(page - 1)
            n = (a.ctypes.data + a.nbytes - lo) & ~(page - 1)
            if n > 0:
                libc = C.CDLL("libc.so.6", use_errno=True)
                libc.madvise(C.c_void_p(lo), C.c_size_t(n), 14)          # MADV_HUGEPAGE
        except Exception:
            pass
    return a


def _ptr(a):
    return C.c_void_p(a.ctypes.data) if a is not None else None


def _f64(a):
    return None if a is None else np.ascontiguousarray(a, dtype=np.float64)


def build_source_tables(pulses, times):
    """Host-side evaluation of what solver.py:176-217 reads from the pulses.

    `pulses`: objects with get_direction/get_row/get_col/get_start_time/get_end_time/magnitude
    (the reference's Pulse API) or with .direction/.row/.col/.start/.end/.table.
    Returns (sources, mag, magz, active, last, first_bad) where first_bad is the
    first call index at which the reference would raise IndexError (a pulse
    still active past the end of its own table)."""
    times = np.asarray(times, dtype=np.float64)
    n_calls = len(times)
    n = len(pulses)
    mag = np.zeros((n, n_calls))
    active = np.zeros((n, n_calls), dtype=np.uint8)
    srcs = (L.Source * max(n, 1))()
    first_bad = None
    for p, pulse in enumerate(pulses):
        if hasattr(pulse, "get_direction"):
            d, r, c = pulse.get_direction(), pulse.get_row(), pulse.get_col()
            t0, t1, table = pulse.get_start_time(), pulse.get_end_time(), pulse._magnitude_vector
        else:
            d, r, c, t0, t1, table = pulse.direction, pulse.row, pulse.col, pulse.start, pulse.end, pulse.table
        table = np.asarray(table, dtype=np.float64)
        m = min(len(table), n_calls)
        mag[p, :m] = table[:m]
        active[p] = (t0 < times) & (times < t1)
        if len(table) < n_calls:
            bad = np.nonzero(active[p, len(table):])[0]
            if len(bad):
                b = len(table) + int(bad[0])
                first_bad = b if first_bad is None else min(first_bad, b)
        srcs[p].kind, srcs[p].row, srcs[p].col = L.DIRECTION_KIND[d], int(r), int(c)
    magz = mag * IMPEDANCE
    last = np.zeros(n_calls)
    for p in range(n):                       # later pulses in list order win (solver.py:178)
        last = np.where(active[p] != 0, mag[p], last)
    return srcs, mag, magz, active, last, first_bad


class Engine:
    def __init__(self, rows, cols=None, *, courant, one_minus_courant=None, dtype="float64", device=0,
                 reflect=(False, False, False, False), row_begin=0, row_end=None, kernel="stream",
                 tile_rows=0, vec=0, block_warps=0, temporal=1, resident=None, frame_mode=None, kstep_variant=None,
                 halo_depth=None):
        import torch
        if not torch.cuda.is_available():
            raise RuntimeError("fdtd_em_solver_b200 needs a CUDA device: the field update exists only as "
                               "sm_100a kernels (no CPU fallback)")
        self.torch = torch
        self.lib = L.load()
        cols = rows if cols is None else cols
        row_end = rows if row_end is None else row_end
        self.rows, self.cols, self.row_begin, self.row_end = int(rows), int(cols), int(row_begin), int(row_end)
        self.owned = self.row_end - self.row_begin
        # "float32x": fields stored as fp32 hi + bf16 lo planes (6 bytes per value), float64 coefficients and arithmetic
        # (include/fdtd2d.h FDTD_F32X) -- the reduced-storage mode that keeps 1e-5 on long runs; one GPU, whole grid
        self.split = str(dtype) == "float32x"
        self.dtype = np.dtype("float64") if self.split else np.dtype(dtype)
        if self.dtype not in (np.dtype("float64"), np.dtype("float32")):
            raise ValueError("dtype must be float64, float32 or float32x")
        cfg = L.Config()
        cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = self.rows, self.cols, self.row_begin, self.row_end
        cfg.pitch = 0
        cfg.dtype = L.F32X if self.split else (L.F64 if self.dtype == np.float64 else L.F32)
        cfg.device = int(device)
        for k in range(4):
            cfg.reflect[k] = 1 if reflect[k] else 0
        cfg.kernel = L.KERNEL_STREAM if kernel == "stream" else L.KERNEL_EXACT
        cfg.tile_rows, cfg.vec, cfg.block_warps = int(tile_rows), int(vec), int(block_warps)
        cfg.courant = float(courant)
        cfg.one_minus_courant = float(1 - courant) if one_minus_courant is None else float(one_minus_courant)
        self.cfg = cfg
        self.pitch = int(self.lib.fdtd_default_pitch(self.cols, cfg.dtype))
        # y-slabs: rows of h/ex/ey stored from each neighbouring slab.  1 = one row, exchanged every step; D >= 2 = deep
        # halos, ONE exchange per temporal-blocking group of up to D-1 steps.  Default: K+1 under temporal blocking (what
        # stepK_edge needs at a slab interface; bit-exact on 2 GPUs since profiles/r2_slab_check_n2.log), else 1.
        slab = self.row_begin > 0 or self.row_end < self.rows
        if halo_depth is None:
            env = os.environ.get("FDTD_HALO_DEPTH", "")
            halo_depth = int(env) if env else (min(int(temporal) + 1, 9) if (slab and temporal > 1 and kernel == "stream"
                                                                             and self.owned >= int(temporal) + 1) else 1)
        self.halo_depth = int(halo_depth)
        self.lrows = int(self.lib.fdtd_local_rows_halo(C.byref(cfg), self.halo_depth))
        self.handle = C.c_void_p()
        L.check(self.lib, None, self.lib.fdtd_create(C.byref(cfg), C.byref(self.handle)))
        if self.halo_depth != 1:
            L.check(self.lib, self.handle, self.lib.fdtd_set_halo_depth(self.handle, self.halo_depth))
        deep = self.halo_depth > 1
        self.halo_top = (self.halo_depth if deep else 1) if self.row_begin > 0 else 0      # stored rows above the slab
        self.halo_bot = self.halo_depth if (deep and self.row_end < self.rows) else 0      # ... and below (deep only)
        self.row_off = self.row_begin - self.halo_top                                      # global row of local row 0
        tdt = torch.float64 if self.dtype == np.float64 else torch.float32
        dev = torch.device("cuda", int(device))
        self.device = dev
        # the ONLY role torch plays: owner of the device buffers
        if self.split:      # six field buffers of hi (4 B) + lo (2 B) planes, two float64 coefficient planes
            self.buffers = [torch.zeros((self.lrows * self.pitch * 6,), dtype=torch.uint8, device=dev) for _ in range(6)] + \
                           [torch.zeros((self.lrows, self.pitch), dtype=torch.float64, device=dev) for _ in range(2)]
        else:
            self.buffers = [torch.zeros((self.lrows, self.pitch), dtype=tdt, device=dev) for _ in range(8)]
        torch.cuda.synchronize(dev)
        L.check(self.lib, self.handle, self.lib.fdtd_bind_buffers(self.handle, *[C.c_void_p(b.data_ptr()) for b in self.buffers]))
        self.first_bad = None
        self._keep = None
        # non-plain tiles of a temporal-blocking group: 3 = stepK_edge (library default), 0 = masked single-step passes
        # through scratch generations
        if frame_mode is None and os.environ.get("FDTD_FRAME_MODE", "") != "":
            frame_mode = int(os.environ["FDTD_FRAME_MODE"])
        if frame_mode is not None:
            L.check(self.lib, self.handle, self.lib.fdtd_set_frame_mode(self.handle, int(frame_mode)))
        # K-step kernel of temporal blocking: 1 = stepK_lean (library default), 0 = stepK_stream
        if kstep_variant is None and os.environ.get("FDTD_KSTEP_VARIANT", "") != "":
            kstep_variant = int(os.environ["FDTD_KSTEP_VARIANT"])
        if kstep_variant is not None:
            L.check(self.lib, self.handle, self.lib.fdtd_set_kstep_variant(self.handle, int(kstep_variant)))
        self.frame_mode, self.kstep_variant = self.tuning()["frame_mode"], self.tuning()["kstep_variant"]
        self.temporal = 1
        if temporal > 1:
            self.enable_temporal_blocking(temporal)
        if resident is None:                # default: only when the caller left the launch strategy to the library
            resident = (_resident_default() and temporal <= 1 and kernel == "stream"
                        and not (tile_rows or vec or block_warps))
        self.set_resident(resident)

    def set_resident(self, on=True, max_cells=0):
        """run() advances all calls up to the next snapshot in ONE cooperative launch (kernel step_resident; step() = a
        launch of one call) when the
        whole grid is on this GPU and (rows+1)*(cols+1) <= max_cells (0 = the library default, 512*512); larger grids
        and slabs keep the per-call / temporal-blocking path.  Same bits either way."""
        L.check(self.lib, self.handle, self.lib.fdtd_set_resident(self.handle, 1 if on else 0, int(max_cells)))
        self.resident = bool(on)

    def resident_active(self):
        """Would run() use the resident kernel for the scene as it is now?"""
        return bool(self.lib.fdtd_resident_active(self.handle))

    def tuning(self):
        """What the library runs with, its own defaults included (fdtd_get_tuning)."""
        out = (C.c_int32 * 8)()
        L.check(self.lib, self.handle, self.lib.fdtd_get_tuning(self.handle, out))
        return dict(zip(("tile_rows", "vec", "block_warps", "temporal", "frame_mode", "kstep_variant", "halo_depth",
                         "generations"), (int(v) for v in out)))

    def enable_temporal_blocking(self, steps_per_pass=2, scratch=None):
        """K leapfrog steps per HBM pass in run().  Scratch generations of h/ex/ey (3 more buffers for K=2, 6 more for
        K=3..8) are only needed by the masked pass chain (frame mode 0; y-slabs with fewer than K+1 halo rows): frame
        mode 3 reads one generation and writes the other.  scratch=None allocates them only where that chain can run
        (a scene stepK_edge cannot take -- a TF/SF "right" line at column 0 -- then steps call by call); True forces."""
        torch = self.torch
        K = int(steps_per_pass)
        slab = self.row_begin > 0 or self.row_end < self.rows
        if scratch is None:
            scratch = self.frame_mode != 3 or (slab and self.halo_depth < K + 1)
        need = 8 if (K <= 1 or not scratch) else (11 if K == 2 else 14)
        while len(self.buffers) < need:
            extra = [torch.zeros((self.lrows, self.pitch), dtype=self.buffers[0].dtype, device=self.device) for _ in range(3)]
            torch.cuda.synchronize(self.device)
            L.check(self.lib, self.handle, self.lib.fdtd_bind_scratch(self.handle, *[C.c_void_p(b.data_ptr()) for b in extra]))
            self.buffers += extra
        L.check(self.lib, self.handle, self.lib.fdtd_set_temporal_blocking(self.handle, K))
        self.temporal = K

    # ---- static inputs -------------------------------------------------------------------
    def upload_coefficients(self, eps_coef, mu_coef):
        e, m = _f64(eps_coef), _f64(mu_coef)
        if e.shape != (self.rows, self.cols) or m.shape != (self.rows, self.cols):
            raise ValueError("coefficient arrays must be (rows, cols)")
        L.check(self.lib, self.handle, self.lib.fdtd_upload_coeffs(self.handle, _ptr(e), _ptr(m), self.cols))

    def paint_coefficients(self, shapes, dt_over_ds, eps_background=EPSILON, mu_background=MU):
        """Device-side MaterialArray painting + coefficient build.  `shapes`: [(kind, (a0..), eps_abs, mu_abs), ...]
        in painting order (see include/fdtd2d.h fdtd_shape)."""
        arr = (L.Shape * max(len(shapes), 1))()
        for k, (kind, a, eps, mu) in enumerate(shapes):
            arr[k].kind = kind
            for q, v in enumerate(a):
                arr[k].a[q] = int(v)
            arr[k].eps, arr[k].mu = float(eps), float(mu)
        L.check(self.lib, self.handle, self.lib.fdtd_paint_coeffs(self.handle, len(shapes), arr, float(eps_background),
                                                                 float(mu_background), float(dt_over_ds)))

    def fill_synthetic_coefficients(self, seed, dt_over_ds):
        L.check(self.lib, self.handle, self.lib.fdtd_fill_synthetic_coeffs(self.handle, int(seed), float(dt_over_ds),
                                                                          EPSILON, MU))

    def set_sources(self, pulses, times):
        srcs, mag, magz, active, last, self.first_bad = build_source_tables(pulses, times)
        self._keep = (srcs, mag, magz, active, last)
        L.check(self.lib, self.handle, self.lib.fdtd_set_sources(
            self.handle, len(pulses), srcs, len(times), _ptr(mag), _ptr(magz), _ptr(active), _ptr(last)))

    def set_reflectors(self, rects):
        a = np.ascontiguousarray(np.asarray(rects, dtype=np.int64).reshape(-1, 4))
        L.check(self.lib, self.handle, self.lib.fdtd_set_reflectors(self.handle, len(a), _ptr(a)))

    def set_probes(self, cells):
        a = np.ascontiguousarray(np.asarray(cells, dtype=np.int64).reshape(-1, 2))
        self.n_probes = len(a)
        L.check(self.lib, self.handle, self.lib.fdtd_set_probes(self.handle, len(a), _ptr(a)))

    # ---- fields ----------------------------------------------------------------------------
    def set_fields(self, h=None, ex=None, ey=None):
        h, ex, ey = _f64(h), _f64(ex), _f64(ey)
        for a, shape in ((h, (self.rows, self.cols)), (ex, (self.rows + 1, self.cols)), (ey, (self.rows, self.cols + 1))):
            if a is not None and a.shape != shape:
                raise ValueError("field shape %s, expected %s" % (a.shape, shape))
        L.check(self.lib, self.handle, self.lib.fdtd_set_fields(self.handle, _ptr(h), _ptr(ex), _ptr(ey)))

    def zero_fields(self):
        L.check(self.lib, self.handle, self.lib.fdtd_zero_fields(self.handle))

    def get_fields(self, which="hxy", out=None):
        """Global-shaped float64 arrays; only the owned rows are written (others stay as given / zero)."""
        out = out or {}
        h = out.get("h", np.zeros((self.rows, self.cols))) if "h" in which else None
        ex = out.get("ex", np.zeros((self.rows + 1, self.cols))) if "x" in which else None
        ey = out.get("ey", np.zeros((self.rows, self.cols + 1))) if "y" in which else None
        L.check(self.lib, self.handle, self.lib.fdtd_get_fields(self.handle, _ptr(h), _ptr(ex), _ptr(ey)))
        return h, ex, ey

    def generation_buffers(self):
        """The torch tensors that currently hold (h, ex, ey): generations 0/1 are buffers[0..5] interleaved as
        h0,h1,ex0,ex1,ey0,ey1; scratch generations 2 and 3 follow the coefficient planes as h,ex,ey triples."""
        g = int(self.lib.fdtd_current_generation(self.handle))
        if g < 2:
            return self.buffers[0 + g], self.buffers[2 + g], self.buffers[4 + g]
        base = 8 + 3 * (g - 2)
        return self.buffers[base], self.buffers[base + 1], self.buffers[base + 2]

    # ---- time stepping -----------------------------------------------------------------------
    def step(self, n):
        if self.first_bad is not None and n >= self.first_bad:
            raise IndexError("index %d is out of bounds for the pulse magnitude table" % n)
        L.check(self.lib, self.handle, self.lib.fdtd_step(self.handle, int(n)))

    def run(self, first_step, n_steps, snapshot_every=0, total_calls=None, alloc=None):
        """Steps [first_step, first_step+n_steps); returns the float64 snapshot stack (or None).  The stack is an
        ordinary numpy array: the library stages each snapshot through its own small pinned ring.  `alloc(shape)`
        may supply the stack instead of np.empty (e.g. a np.lib.format.open_memmap, so a long run streams straight
        into its .npy file); it must return a writeable C-contiguous float64 array of that shape."""
        first_step, n_steps = int(first_step), int(n_steps)
        total_calls = first_step + n_steps if total_calls is None else int(total_calls)
        bad = None
        if self.first_bad is not None and first_step + n_steps > self.first_bad:
            bad = self.first_bad
            n_steps = max(0, bad - first_step)
        snaps, n_snaps = None, 0
        if snapshot_every:
            k = int(snapshot_every)
            n_snaps = (first_step + n_steps) // k - first_step // k
            if n_snaps:
                shape = (n_snaps, self.owned, self.cols)
                snaps = empty_hugepages(shape) if alloc is None else alloc(shape)
                if (snaps.shape != shape or snaps.dtype != np.float64 or not snaps.flags.c_contiguous
                        or not snaps.flags.writeable):
                    raise ValueError("alloc() must return a writeable C-contiguous float64 array of shape %r" % (shape,))
        got = C.c_int64(0)
        L.check(self.lib, self.handle, self.lib.fdtd_run(
            self.handle, first_step, n_steps, int(snapshot_every), _ptr(snaps) if snaps is not None else None,
            n_snaps, total_calls, C.byref(got)))
        if bad is not None:
            raise IndexError("index %d is out of bounds for the pulse magnitude table" % bad)
        return snaps

    def run_streamed(self, eps_coef, mu_coef, first_step, n_steps, zero_fields=True, want_h=True, h_out=None):
        """upload_coefficients + zero_fields + run + get_fields("h") as one pipeline over row bands (fdtd_run_streamed):
        the coefficient rows travel up while the calls already run on the bands above, h travels back behind each band's
        last call.  Returns h (global shape, owned rows written) or None.  Pinned host arrays (torch pin_memory, or
        cudaHostRegister-ed numpy) make the copies asynchronous; ordinary numpy arrays work without the overlap."""
        e, m = _f64(eps_coef), _f64(mu_coef)
        if e.shape != (self.rows, self.cols) or m.shape != (self.rows, self.cols):
            raise ValueError("coefficient arrays must be (rows, cols)")
        first_step, n_steps = int(first_step), int(n_steps)
        bad = None
        if self.first_bad is not None and first_step + n_steps > self.first_bad:
            bad = self.first_bad
            n_steps = max(0, bad - first_step)
        h = None
        if want_h:
            h = np.zeros((self.rows, self.cols)) if h_out is None else h_out
        L.check(self.lib, self.handle, self.lib.fdtd_run_streamed(self.handle, _ptr(e), _ptr(m), self.cols, 1 if zero_fields else 0,
                                                                  first_step, n_steps, _ptr(h), self.cols))
        if bad is not None:
            raise IndexError("index %d is out of bounds for the pulse magnitude table" % bad)
        return h

    # ---- realtime view --------------------------------------------------------------
    def run_enqueue(self, first_step, n_steps):
        """Queue calls [first_step, first_step+n_steps) and return without waiting (fdtd_run_enqueue).  Returns the number
        of calls queued: a pulse that outlives its table (App. A-Q10) ends the batch early, and the next batch raises
        the reference's IndexError."""
        first_step, n_steps = int(first_step), int(n_steps)
        if self.first_bad is not None and first_step + n_steps > self.first_bad:
            if first_step >= self.first_bad:
                raise IndexError("index %d is out of bounds for the pulse magnitude table" % self.first_bad)
            n_steps = self.first_bad - first_step
        L.check(self.lib, self.handle, self.lib.fdtd_run_enqueue(self.handle, first_step, n_steps))
        return n_steps

    def view_shape(self, row_stride=1, col_stride=1):
        first = -(-self.row_begin // row_stride) * row_stride
        rows = (self.row_end - 1 - first) // row_stride + 1 if first < self.row_end else 0
        return rows, (self.cols - 1) // col_stride + 1

    def view_begin(self, row_stride=1, col_stride=1):
        """Queue a picture of h[::row_stride, ::col_stride] (this slab's rows of it) behind everything queued so far; its
        device-to-host copy runs on the side stream, next to whatever is queued after it."""
        L.check(self.lib, self.handle, self.lib.fdtd_view_begin(self.handle, int(row_stride), int(col_stride)))
        self._views = getattr(self, "_views", [])
        self._views.append(self.view_shape(int(row_stride), int(col_stride)))

    def view_end(self):
        """The oldest picture in flight, as a float64 array (waits for its copy only, not for later work)."""
        shape = self._views.pop(0)
        out = np.empty(shape, dtype=np.float64)
        rows, cols = C.c_int64(0), C.c_int64(0)
        L.check(self.lib, self.handle, self.lib.fdtd_view_end(self.handle, _ptr(out) if out.size else _ptr(np.empty(1)),
                                                              max(out.size, 1), C.byref(rows), C.byref(cols)))
        assert (rows.value, cols.value) == tuple(shape)
        return out

    def sync(self):
        L.check(self.lib, self.handle, self.lib.fdtd_sync(self.handle))

    def probes(self):
        n = C.c_int64(0)
        L.check(self.lib, self.handle, self.lib.fdtd_probe_fetch(self.handle, None, 0, C.byref(n)))
        out = np.zeros((n.value, getattr(self, "n_probes", 0), 3))
        if out.size:
            L.check(self.lib, self.handle, self.lib.fdtd_probe_fetch(self.handle, _ptr(out), n.value, C.byref(n)))
        return out

    def clear_probes(self):
        L.check(self.lib, self.handle, self.lib.fdtd_probe_clear(self.handle))

    def kernel_timing(self, enable=True):
        """Bracket every launch of the plain K-step kernel with CUDA events on its stream (measurement hook)."""
        L.check(self.lib, self.handle, self.lib.fdtd_kernel_timing(self.handle, 1 if enable else 0))

    def kernel_timing_fetch(self):
        n, tot, lo, hi, cu = C.c_int32(0), C.c_double(0), C.c_double(0), C.c_double(0), C.c_int64(0)
        L.check(self.lib, self.handle, self.lib.fdtd_kernel_timing_fetch(self.handle, C.byref(n), C.byref(tot), C.byref(lo),
                                                                        C.byref(hi), C.byref(cu)))
        return dict(launches=n.value, total_ms=tot.value, min_ms=lo.value, max_ms=hi.value, cell_updates=cu.value)

    def uniform_mu(self):
        """True when the library found the mu_coef plane uniform and runs the K-step kernel without that stream."""
        return int(self.lib.fdtd_uniform_mu(self.handle)) == 1

    def rescan_coefficients(self):
        """After writing self.buffers[6] / [7] (the coefficient planes) directly."""
        L.check(self.lib, self.handle, self.lib.fdtd_rescan_coeffs(self.handle))

    def stats(self):
        s = L.Stats()
        L.check(self.lib, self.handle, self.lib.fdtd_get_stats(self.handle, C.byref(s)))
        return dict(steps=s.steps, cell_updates=s.cell_updates, kernel_launches=s.kernel_launches,
                    halo_bytes_sent=s.halo_bytes_sent, last_run_ms=s.last_run_ms)

    # ---- multi-GPU ---------------------------------------------------------------------------
    def comm_init(self, rank, nranks, unique_id):
        buf = (C.c_char * 128).from_buffer_copy(bytes(unique_id))
        L.check(self.lib, self.handle, self.lib.fdtd_comm_init(self.handle, C.cast(buf, C.c_void_p), rank, nranks))

    @staticmethod
    def comm_unique_id():
        """The 128-byte NCCL id rank 0 creates and every rank passes to comm_init()."""
        return comm_unique_id()

    def halo_exchange(self):
        L.check(self.lib, self.handle, self.lib.fdtd_halo_exchange(self.handle))

    def close(self):
        if getattr(self, "handle", None) is not None and self.handle:
            self.lib.fdtd_destroy(self.handle)
            self.handle = None
        self.buffers = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


def comm_unique_id():
    lib = L.load()
    buf = (C.c_char * 128)()
    L.check(lib, None, lib.fdtd_comm_unique_id(C.cast(buf, C.c_void_p)))
    return bytes(buf)
