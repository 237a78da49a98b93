"""TEST INFRASTRUCTURE ONLY -- ctypes driver of tests/host_emul/stream_emul.cpp.

Builds (g++, no nvcc needed) the SIMT-on-CPU emulation of the streaming kernels of fdtd_em_solver_b200/csrc/kernels.cuh
(tests/host_emul/simt.h: the device code itself, one warp at a time, every lane a fiber) and feeds it oracle scenes.
Used by tests/test_stream_emulation.py; nothing in the package imports it."""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

from resident_emul import MAX_RECT, MAX_SRC, clip_rect

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "host_emul", "stream_emul.cpp")
OUT_DIR = os.path.join(HERE, "host_emul", "_build")
OUT = os.path.join(OUT_DIR, "libstream_emul.so")
CSRC = os.path.join(ROOT, "fdtd_em_solver_b200", "csrc")
CUDA_INC = "/usr/local/cuda/include"


class EmuSrc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("row", C.c_int32), ("col", C.c_int32), ("pad", C.c_int32),
                ("mag", C.c_double), ("magz", C.c_double)]


class StreamScene(C.Structure):
    _fields_ = [("a", C.c_void_p * 3), ("b", C.c_void_p * 3), ("muc", C.c_void_p), ("epc", C.c_void_p),
                ("pitch", C.c_int64), ("NR", C.c_int32), ("NC", C.c_int32), ("absorb", C.c_int32 * 4),
                ("S", C.c_double), ("oms", C.c_double), ("last_mag", C.c_double),
                ("n_src", C.c_int32), ("src", EmuSrc * MAX_SRC), ("n_rect", C.c_int32),
                ("rh", (C.c_int32 * 4) * MAX_RECT), ("rex", (C.c_int32 * 4) * MAX_RECT), ("rey", (C.c_int32 * 4) * MAX_RECT)]


def available():
    return shutil.which("g++") is not None and os.path.exists(os.path.join(CUDA_INC, "cuda_runtime.h"))


def build():
    deps = [SRC, os.path.join(HERE, "host_emul", "simt.h")] + [os.path.join(CSRC, f) for f in ("kernels.cuh", "edge.cuh", "resident.cuh")]
    if os.path.isfile(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in deps):
        return OUT
    os.makedirs(OUT_DIR, exist_ok=True)
    cmd = ["g++", "-std=c++20", "-O1", "-ffp-contract=off", "-fPIC", "-shared", "-I", CUDA_INC, "-I", CSRC,
           "-I", os.path.join(ROOT, "include"), "-I", os.path.join(HERE, "host_emul"), SRC, "-o", OUT + ".tmp"]
    res = subprocess.run(cmd, capture_output=True, text=True)
    if res.returncode != 0:
        raise RuntimeError("g++ failed:\n" + res.stdout + res.stderr)
    os.replace(OUT + ".tmp", OUT)
    return OUT


_lib = None


def load():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.emul_step_stream.restype = C.c_int
        _lib.emul_step_stream.argtypes = [C.POINTER(StreamScene)] + [C.c_int] * 5
        _lib.emul_stepK_stream.restype = C.c_int
        _lib.emul_stepK_stream.argtypes = [C.POINTER(StreamScene)] + [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _lib.emul_stepK_lean_split.restype = C.c_int
        _lib.emul_stepK_lean_split.argtypes = [C.POINTER(StreamScene)] + [C.c_int] * 5 + [C.c_void_p, C.c_int, C.c_int,
                                                                                         C.c_longlong]
        _lib.emul_stepK_edge.restype = C.c_int
        _lib.emul_stepK_edge.argtypes = [C.POINTER(StreamScene), C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p,
                                         C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_void_p, C.c_longlong]
    return _lib


class Scene:
    """Generation A = the oracle-style fields `f`; generation B pre-filled with `fill`.  The pulse record is the one of
    call `n` (what fdtd2d.cu fill_params puts into StepParams for a single-step launch)."""

    def __init__(self, f, setup, times, n, fill=np.nan, pad=0.0):
        from fdtd_em_solver_b200.engine import build_source_tables
        NR, NC = f.h.shape
        self.NR, self.NC = NR, NC
        pitch = ((NC + 1 + 31) // 32) * 32

        def plane(a, value=0.0):
            p = np.full((NR + 1, pitch), value)
            if a is not None:
                p[:a.shape[0], :a.shape[1]] = a
            return p
        self.a = [plane(f.h, pad), plane(f.ex, pad), plane(f.ey, pad)]          # pad: what lies outside the arrays
        self.b = [plane(None, fill) for _ in range(3)]
        self.muc, self.epc = plane(setup.mu_coef, pad), plane(setup.eps_coef, pad)
        sc = StreamScene()
        for k in range(3):
            sc.a[k], sc.b[k] = self.a[k].ctypes.data, self.b[k].ctypes.data
        sc.muc, sc.epc, sc.pitch, sc.NR, sc.NC = self.muc.ctypes.data, self.epc.ctypes.data, pitch, NR, NC
        for k, reflect in enumerate(setup.reflect_walls):
            sc.absorb[k] = 0 if reflect else 1
        S = setup.courant
        sc.S, sc.oms = S, 1 - S
        srcs, mag, magz, active, last, _ = build_source_tables(setup.pulses, times)
        k = 0
        for p in range(len(setup.pulses)):
            if active[p, n]:
                assert k < MAX_SRC
                q = sc.src[k]
                q.kind, q.row, q.col, q.mag, q.magz = srcs[p].kind, srcs[p].row, srcs[p].col, mag[p, n], magz[p, n]
                k += 1
        sc.n_src = k
        sc.last_mag = float(last[n]) if len(setup.pulses) else 0.0
        sc.n_rect = len(setup.reflectors)
        for r, rect in enumerate(setup.reflectors):
            for arr, shape in ((sc.rh, (NR, NC)), (sc.rex, (NR + 1, NC)), (sc.rey, (NR, NC + 1))):
                for q, v in enumerate(clip_rect(rect, *shape)):
                    arr[r][q] = v
        self.sc = sc

    def result(self):
        NR, NC = self.NR, self.NC
        return self.b[0][:NR, :NC].copy(), self.b[1][:NR + 1, :NC].copy(), self.b[2][:NR, :NC + 1].copy()


def step_stream(f, setup, times, n, vec=2, tile_rows=8, warps=1):
    """One launch of step_stream<double, vec> over the whole index space (call n)."""
    s = Scene(f, setup, times, n)
    rc = load().emul_step_stream(C.byref(s.sc), vec, tile_rows, warps, 0, s.NR + 1)
    assert rc == 0
    return s.result()


def stepK_stream(f, setup, times, n, K, tile_rows, warps, row_lo, row_hi, skip, variant=0):
    """One launch of stepK_stream<double, 2, K> (variant 1: stepK_lean) with the skip map of the frame plan."""
    s = Scene(f, setup, times, n)
    skip = np.ascontiguousarray(skip, dtype=np.uint8)
    rc = load().emul_stepK_stream(C.byref(s.sc), K, tile_rows, warps, row_lo, row_hi, skip.ctypes.data,
                                  skip.shape[1], skip.shape[0], variant)
    assert rc == 0
    return s.result()


def _split_planes(a, pitch, rows, hi_fill):
    """A field as the library stores it under FDTD_F32X: (rows x pitch) float32 hi plane, then the uint16 lo plane."""
    hi = np.full((rows, pitch), hi_fill, dtype=np.float32)
    lo = np.zeros((rows, pitch), dtype=np.uint16)
    if a is not None:
        x = np.asarray(a, dtype=np.float64)
        h32 = x.astype(np.float32)
        r32 = (x - h32.astype(np.float64)).astype(np.float32)
        u = r32.view(np.uint32).astype(np.uint64)
        bits = (np.where((u & 0x7F800000) != 0x7F800000, u + 0x7FFF + ((u >> 16) & 1), u) >> 16).astype(np.uint16)
        hi[:x.shape[0], :x.shape[1]] = h32
        lo[:x.shape[0], :x.shape[1]] = bits
    buf = np.empty(rows * pitch * 6, dtype=np.uint8)
    buf[:rows * pitch * 4] = hi.view(np.uint8).ravel()
    buf[rows * pitch * 4:] = lo.view(np.uint8).ravel()
    return buf


def _split_values(buf, pitch, rows):
    hi = buf[:rows * pitch * 4].view(np.float32).reshape(rows, pitch)
    lo = buf[rows * pitch * 4:].view(np.uint16).reshape(rows, pitch)
    return hi.astype(np.float64) + (lo.astype(np.uint32) << 16).view(np.float32).astype(np.float64)


def stepK_lean_split(f, setup, times, n, K, tile_rows, warps, row_lo, row_hi, skip):
    """One launch of stepK_lean_split<K>: the plain tiles of the frame plan under split field storage.  f holds the fields as
    storage keeps them; returns generation B decoded (cells no tile stored are NaN)."""
    s = Scene(f, setup, times, n)
    skip = np.ascontiguousarray(skip, dtype=np.uint8)
    rows, pitch = s.NR + 1, s.sc.pitch
    a_planes = [_split_planes(x, pitch, rows, np.nan) for x in (f.h, f.ex, f.ey)]
    b_planes = [_split_planes(None, pitch, rows, np.nan) for _ in range(3)]
    for k in range(3):
        s.sc.a[k], s.sc.b[k] = a_planes[k].ctypes.data, b_planes[k].ctypes.data
    rc = load().emul_stepK_lean_split(C.byref(s.sc), K, tile_rows, warps, row_lo, row_hi, skip.ctypes.data,
                                      skip.shape[1], skip.shape[0], C.c_longlong(rows * pitch * 4))
    assert rc == 0
    vals = [_split_values(b, pitch, rows) for b in b_planes]
    NR, NC = s.NR, s.NC
    return vals[0][:NR, :NC].copy(), vals[1][:NR + 1, :NC].copy(), vals[2][:NR, :NC + 1].copy()


def stepK_edge(f, setup, times, n, K, tiles, probes=(), fill=np.nan, split=False):
    """One launch of stepK_edge<double, 2, K>: calls n .. n+K-1 for every listed tile (tx, ta, tb[, class]): a tile with a class
    runs through that instantiation of the kernel (edge.cuh EDGE_CLASSES), the others through the all-features one.  Returns the fields
    of generation B (cells no tile owns keep `fill`) and the probe records (K, n_probe, 3)."""
    from fdtd_em_solver_b200.engine import build_source_tables
    s = Scene(f, setup, times, n, fill=fill, pad=np.nan)          # NaN outside the arrays: nothing read there may be kept
    s.sc.n_src = 0
    srcs, mag, magz, active, last, _ = build_source_tables(setup.pulses, times)
    stride = len(setup.pulses)
    count = np.zeros(K, dtype=np.int32)
    lastv = np.zeros(K)
    recs = (EmuSrc * max(K * stride, 1))()
    for m in range(K):
        k = 0
        for p in range(stride):
            if active[p, n + m]:
                q = recs[m * stride + k]
                q.kind, q.row, q.col, q.mag, q.magz = srcs[p].kind, srcs[p].row, srcs[p].col, mag[p, n + m], magz[p, n + m]
                k += 1
        count[m] = k
        lastv[m] = last[n + m] if stride else 0.0
    tl = np.ascontiguousarray([list(t) + [0] * (4 - len(t)) for t in tiles], dtype=np.int32).reshape(-1, 4)
    pij = np.ascontiguousarray(np.asarray(probes, dtype=np.int32).reshape(-1, 2))
    pout = np.full((K, max(len(pij), 1), 3), np.nan)
    lo_off = 0
    if split:
        # split=True: generation A holds the fields as split storage keeps them, generation B (hi planes NaN) receives the
        # encoded results; the coefficient planes stay float64
        rows, pitch = s.NR + 1, s.sc.pitch
        lo_off = rows * pitch * 4
        a_planes = [_split_planes(x, pitch, rows, np.nan) for x in (f.h, f.ex, f.ey)]
        b_planes = [_split_planes(None, pitch, rows, fill) for _ in range(3)]
        for k in range(3):
            s.sc.a[k], s.sc.b[k] = a_planes[k].ctypes.data, b_planes[k].ctypes.data
    rc = load().emul_stepK_edge(C.byref(s.sc), K, tl.ctypes.data, len(tl), count.ctypes.data, lastv.ctypes.data,
                                C.cast(recs, C.c_void_p), stride, len(pij), pij.ctypes.data, pout.ctypes.data, lo_off)
    assert rc == 0
    if split:
        NR, NC = s.NR, s.NC
        vals = [_split_values(b, pitch, rows) for b in b_planes]
        return (vals[0][:NR, :NC].copy(), vals[1][:NR + 1, :NC].copy(), vals[2][:NR, :NC + 1].copy()), pout[:, :len(pij)]
    return s.result(), pout[:, :len(pij)]
