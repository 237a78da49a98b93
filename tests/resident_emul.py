"""TEST INFRASTRUCTURE ONLY -- ctypes driver of tests/host_emul/resident_emul.cu and window_emul.cu.

Builds the CPU emulations of the window algorithms of step_resident (fdtd_em_solver_b200/csrc/resident.cuh) and
frame_window (csrc/window.cuh) -- the kernels' own __host__ __device__ phase functions, run by plain loops -- and feeds
them oracle scenes.  Used by tests/test_resident_emulation.py and tests/test_window_emulation.py; nothing in the
package imports it.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_emul", "resident_emul.cu")
OUT_DIR = os.path.join(HERE, "host_emul", "_build")
OUT = os.path.join(OUT_DIR, "libresident_emul.so")
CSRC = os.path.join(ROOT, "fdtd_em_solver_b200", "csrc")

MAX_RECT, MAX_PROBE, MAX_SRC = 24, 16, 16


class EmuSource(C.Structure):
    _fields_ = [("kind", C.c_int32), ("row", C.c_int32), ("col", C.c_int32), ("pad", C.c_int32),
                ("mag", C.c_double), ("magz", C.c_double)]


class EmuScene(C.Structure):
    _fields_ = [("gen0", C.c_void_p * 3), ("gen1", C.c_void_p * 3), ("muc", C.c_void_p), ("epc", C.c_void_p),
                ("pitch", C.c_int64), ("NR", C.c_int32), ("NC", C.c_int32), ("absorb", C.c_int32 * 4),
                ("S", C.c_double), ("oms", C.c_double), ("n_rect", C.c_int32),
                ("rh", (C.c_int32 * 4) * MAX_RECT), ("rex", (C.c_int32 * 4) * MAX_RECT), ("rey", (C.c_int32 * 4) * MAX_RECT),
                ("n_probe", C.c_int32), ("probe_ij", (C.c_int32 * 2) * MAX_PROBE), ("probe_out", C.c_void_p),
                ("count", C.c_void_p), ("last", C.c_void_p), ("src", C.c_void_p), ("stride", C.c_int32)]


def available():
    return (shutil.which("nvcc") or os.path.exists("/usr/local/cuda/bin/nvcc")) is not None


def _build(src, out):
    deps = [src] + [os.path.join(CSRC, f) for f in ("window.cuh", "resident.cuh", "kernels.cuh")]
    if os.path.isfile(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    os.makedirs(OUT_DIR, exist_ok=True)
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    cmd = [nvcc, "-std=c++17", "-O2", "-gencode", "arch=compute_100a,code=sm_100a", "-fmad=false",
           "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared", "-I", CSRC, "-I", os.path.join(ROOT, "include"),
           src, "-o", out + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("nvcc failed:\n" + res.stdout + res.stderr)
    os.replace(out + ".tmp", out)
    return out


def build():
    return _build(SRC, OUT)


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.resident_emulate_f64.restype = C.c_int
        _lib.resident_emulate_f64.argtypes = [C.POINTER(EmuScene), C.c_int, C.c_int, C.c_longlong, C.c_int]
    return _lib


def clip_rect(rect, nrows, ncols):
    """Python slice semantics of a[r0:r1, c0:c1] on an (nrows, ncols) array."""
    r0, r1, _ = slice(rect[0], rect[1]).indices(nrows)
    c0, c1, _ = slice(rect[2], rect[3]).indices(ncols)
    return r0, r1, c0, c1


def source_records(pulses, times):
    """Per-call records of the active pulses in list order: what fdtd2d.cu resident_tables uploads.  Built from the
    PRODUCT's table builder (engine.build_source_tables)."""
    from fdtd_em_solver_b200.engine import build_source_tables
    srcs, mag, magz, active, last, first_bad = build_source_tables(pulses, times)
    n, L = len(pulses), len(times)
    stride = max(n, 1)
    count = np.zeros(L, dtype=np.int32)
    rec = (EmuSource * (L * stride))()
    for t in range(L):
        k = 0
        for p in range(n):
            if not active[p, t]:
                continue
            q = rec[t * stride + k]
            q.kind, q.row, q.col, q.mag, q.magz = srcs[p].kind, srcs[p].row, srcs[p].col, mag[p, t], magz[p, t]
            k += 1
        count[t] = k
    return count, np.ascontiguousarray(last, dtype=np.float64), rec, stride, first_bad


def emulate(f, setup, times, first_step, n_steps, geom=0, order=0, probe_cells=()):
    """Advance oracle-style fields `f` (h, ex, ey attributes; not modified) by calls [first_step, first_step+n_steps)
    with the emulated resident kernel.  Returns (h, ex, ey, probe_records)."""
    lib = load()
    NR, NC = f.h.shape
    pitch = ((NC + 1 + 31) // 32) * 32                       # fdtd_default_pitch

    def plane(a):
        p = np.zeros((NR + 1, pitch))
        if a is not None:
            p[:a.shape[0], :a.shape[1]] = a
        return p
    g0 = [plane(f.h), plane(f.ex), plane(f.ey)]
    g1 = [plane(None) for _ in range(3)]
    muc, epc = plane(setup.mu_coef), plane(setup.eps_coef)
    sc = EmuScene()
    for k in range(3):
        sc.gen0[k] = g0[k].ctypes.data
        sc.gen1[k] = g1[k].ctypes.data
    sc.muc, sc.epc, sc.pitch, sc.NR, sc.NC = muc.ctypes.data, epc.ctypes.data, pitch, NR, NC
    for k, reflect in enumerate(setup.reflect_walls):
        sc.absorb[k] = 0 if reflect else 1
    S = setup.courant
    sc.S, sc.oms = S, 1 - S
    sc.n_rect = len(setup.reflectors)
    for r, rect in enumerate(setup.reflectors):
        for arr, shape in ((sc.rh, (NR, NC)), (sc.rex, (NR + 1, NC)), (sc.rey, (NR, NC + 1))):
            for q, v in enumerate(clip_rect(rect, *shape)):
                arr[r][q] = v
    sc.n_probe = len(probe_cells)
    for k, (i, j) in enumerate(probe_cells):
        sc.probe_ij[k][0], sc.probe_ij[k][1] = i, j
    probe_out = np.zeros((max(n_steps, 1), max(len(probe_cells), 1), 3))
    sc.probe_out = probe_out.ctypes.data if probe_cells else None
    count, last, rec, stride, _ = source_records(setup.pulses, times)
    if len(setup.pulses):
        sc.count, sc.last, sc.src, sc.stride = count.ctypes.data, last.ctypes.data, C.addressof(rec), stride
    g = lib.resident_emulate_f64(C.byref(sc), geom, order, first_step, n_steps)
    assert g in (0, 1), "emulation failed"
    out = g1 if g else g0
    return (out[0][:NR, :NC].copy(), out[1][:NR + 1, :NC].copy(), out[2][:NR, :NC + 1].copy(),
            probe_out[:n_steps, :len(probe_cells)])


# ---- frame_window (csrc/window.cuh) ------------------------------------------------------------------------
WIN_SRC = os.path.join(HERE, "host_emul", "window_emul.cu")
WIN_OUT = os.path.join(OUT_DIR, "libwindow_emul.so")


class WinEmuScene(C.Structure):
    _fields_ = [("gen_a", C.c_void_p * 3), ("gen_b", C.c_void_p * 3), ("muc", C.c_void_p), ("epc", C.c_void_p),
                ("pitch", C.c_int64), ("NR", C.c_int32), ("NC", C.c_int32), ("row_off", C.c_int32), ("lrows", C.c_int32),
                ("row_lo", C.c_int32), ("row_hi", C.c_int32), ("absorb", C.c_int32 * 4),
                ("S", C.c_double), ("oms", C.c_double), ("n_rect", C.c_int32),
                ("rh", (C.c_int32 * 4) * MAX_RECT), ("rex", (C.c_int32 * 4) * MAX_RECT), ("rey", (C.c_int32 * 4) * MAX_RECT),
                ("n_probe", C.c_int32), ("probe_ij", (C.c_int32 * 2) * MAX_PROBE), ("probe_out", C.c_void_p),
                ("count", C.c_void_p), ("last", C.c_void_p), ("src", C.c_void_p), ("stride", C.c_int32)]


_win = None


def load_window():
    global _win
    if _win is None:
        _win = C.CDLL(_build(WIN_SRC, WIN_OUT))
        _win.window_emulate_f64.restype = C.c_int
        _win.window_emulate_f64.argtypes = [C.POINTER(WinEmuScene), C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                            C.c_int, C.c_int, C.c_longlong, C.c_int]
    return _win


def window_emulate(f, setup, times, first_step, K, tiles, tile_rows, tile_cols, order=0, probe_cells=(),
                   apron=0, row_begin=0, row_end=None, fill=np.nan, poison_halo=False, shrink=-1, halo_depth=1):
    """Advance the frame tiles `tiles` ([(tx, ty), ...] of the (tile_rows x tile_cols) grid that starts at row
    `row_begin`) by K calls with the emulated frame_window kernel, reading the oracle-style fields `f` as generation A.
    With row_begin/row_end the context stores only that y-slab (one halo row above, the ex row below), like the library.
    shrink: cells the active box loses per call (-1 = the kernel's rule, window.cuh window_margin).
    poison_halo: the halo rows of generation A (which the library fills by halo exchange) hold NaN instead.
    halo_depth D >= 2: deep halos (fdtd_set_halo_depth) -- the context stores D valid rows of each neighbouring slab; every
    other row it does not own holds NaN.
    Returns (h, ex, ey, probe_records) of generation B as GLOBAL arrays pre-filled with `fill`: cells the kernel did not
    store keep it."""
    lib = load_window()
    NR, NC = f.h.shape
    row_end = NR if row_end is None else row_end
    pitch = ((NC + 1 + 31) // 32) * 32
    deep = halo_depth > 1
    ht = (halo_depth if row_begin > 0 else 0) if deep else (1 if row_begin > 0 else 0)
    hb = halo_depth if (deep and row_end < NR) else 0
    row_off = row_begin - ht
    lrows = (row_end - row_begin) + ht + hb + 1                          # fdtd_local_rows / fdtd_local_rows_halo

    def plane(a, value=0.0):
        p = np.full((lrows, pitch), value)
        if a is not None:
            n = min(a.shape[0], row_off + lrows) - row_off
            p[:n, :a.shape[1]] = a[row_off:row_off + n]
        return p
    ga = [plane(f.h), plane(f.ex), plane(f.ey)]
    if row_end < NR and not deep:
        ga[0][-1, :] = 0.0                                   # the last local row holds ONLY ex[row_end]
        ga[2][-1, :] = 0.0
    if row_end < NR and deep:
        for a in ga:
            a[-1, :] = np.nan                                # row_end + D is stored but never exchanged
    if poison_halo:
        for a in ga:
            if row_begin > 0:
                a[0, :] = np.nan
            if row_end < NR:
                a[-1, :] = np.nan
    gb = [plane(None, fill) for _ in range(3)]
    muc, epc = plane(setup.mu_coef), plane(setup.eps_coef)
    sc = WinEmuScene()
    for k in range(3):
        sc.gen_a[k] = ga[k].ctypes.data
        sc.gen_b[k] = gb[k].ctypes.data
    sc.muc, sc.epc, sc.pitch, sc.NR, sc.NC = muc.ctypes.data, epc.ctypes.data, pitch, NR, NC
    sc.row_off, sc.lrows = row_off, lrows
    sc.row_lo, sc.row_hi = row_begin, row_end + (1 if row_end == NR else 0)
    for k, reflect in enumerate(setup.reflect_walls):
        sc.absorb[k] = 0 if reflect else 1
    S = setup.courant
    sc.S, sc.oms = S, 1 - S
    sc.n_rect = len(setup.reflectors)
    for r, rect in enumerate(setup.reflectors):
        for arr, shape in ((sc.rh, (NR, NC)), (sc.rex, (NR + 1, NC)), (sc.rey, (NR, NC + 1))):
            for q, v in enumerate(clip_rect(rect, *shape)):
                arr[r][q] = v
    sc.n_probe = len(probe_cells)
    for k, (i, j) in enumerate(probe_cells):
        sc.probe_ij[k][0], sc.probe_ij[k][1] = i, j
    probe_out = np.full((K, max(len(probe_cells), 1), 3), np.nan)
    sc.probe_out = probe_out.ctypes.data if probe_cells else None
    count, last, rec, stride, _ = source_records(setup.pulses, times)
    if len(setup.pulses):
        sc.count, sc.last, sc.src, sc.stride = count.ctypes.data, last.ctypes.data, C.addressof(rec), stride
    t = np.ascontiguousarray(np.asarray(tiles, dtype=np.int32).reshape(-1, 2))
    rc = lib.window_emulate_f64(C.byref(sc), len(t), t.ctypes.data, tile_rows, tile_cols, K, apron, order, first_step,
                                shrink)
    assert rc == 0, "emulation failed"

    def globalise(p, rows, cols):
        out = np.full((rows, cols), fill)
        n = min(rows, row_off + lrows) - row_off
        out[row_off:row_off + n] = p[:n, :cols]
        return out
    return (globalise(gb[0], NR, NC), globalise(gb[1], NR + 1, NC), globalise(gb[2], NR, NC + 1),
            probe_out[:, :len(probe_cells)])
