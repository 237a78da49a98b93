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
    deps = [src] + [os.path.join(CSRC, f) for f in ("resident.cuh", "kernels.cuh")]
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
