"""TEST INFRASTRUCTURE ONLY -- ctypes front of oracle/restated.c, call-compatible with oracle.restated.leapfrog.

`build()` compiles the C restatement with gcc into oracle/_build/ (git-ignored); `leapfrog(f, s, time)` takes the same
Fields / Setup objects as the numpy restatement and must leave them in the same state, bit for bit
(tests/test_c_oracle.py).  The gating, the stale `magnitude` and the probe bookkeeping stay in Python, exactly as in
restated.leapfrog; only the array sweeps of solver.py:182-242 run in C.  Roughly 5x faster than numpy on one core
(no temporaries), more with OpenMP; that is its only purpose -- bigger parity cases in the same test budget.
"""
import ctypes as C
import os
import shutil
import subprocess

import numpy as np

from . import restated as R

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "restated.c")
OUT_DIR = os.path.join(HERE, "_build")
LIB = os.path.join(OUT_DIR, "liboracle.so")

KIND = {None: 0, "right": 1, "left": 2, "up": 3, "down": 4}


class Src(C.Structure):
    _fields_ = [("kind", C.c_int), ("row", C.c_long), ("col", C.c_long), ("mag", C.c_double), ("magz", C.c_double)]


def available():
    return shutil.which("gcc") is not None


def build(force=False, openmp=True):
    """Compile restated.c if missing or older than its source.  Returns the path of the shared object."""
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    if not available():
        raise RuntimeError("gcc not found: the C oracle cannot be built")
    os.makedirs(OUT_DIR, exist_ok=True)
    base = ["gcc", "-O2", "-ffp-contract=off", "-fPIC", "-shared", SRC, "-o", LIB]
    if openmp and subprocess.run(base + ["-fopenmp"], capture_output=True).returncode == 0:
        return LIB
    subprocess.run(base, check=True)
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        _lib.oracle_step.restype = None
        _lib.oracle_step.argtypes = [C.c_long, C.c_long] + [C.c_void_p] * 6 + [C.c_double, C.c_double, C.c_void_p,
                                                                              C.c_int, C.c_void_p, C.c_double,
                                                                              C.c_int, C.c_void_p]
    return _lib


def _ptr(a):
    assert a.dtype == np.float64 and a.flags.c_contiguous
    return a.ctypes.data


def leapfrog(f, s, time):
    """One call of Solver.update(time) (solver.py:169-257) through the C restatement.  Mutates `f`."""
    n = f.step
    rows, cols = f.h.shape
    live = [p for p in s.pulses if p.start < time < p.end]              # :177 strict
    srcs = (Src * max(1, len(live)))()
    last = 0.0
    for k, p in enumerate(live):
        last = p.table[n]                                               # IndexError past the table, like the reference
        row, col = p.row, p.col
        if p.direction is None:                                         # python indexing wraps negatives
            row, col = row % rows, col % cols
        assert row >= 0 and col >= 0
        srcs[k] = Src(KIND[p.direction], row, col, last, last * R.IMPEDANCE)
    rect = np.asarray(s.reflectors, dtype=np.int64).reshape(-1, 4)
    rect = np.ascontiguousarray(rect).astype(np.dtype("l"))
    walls = (C.c_int * 4)(*[int(bool(w)) for w in s.reflect_walls])
    hn = np.empty_like(f.h)                                             # fresh: saved snapshots alias the old one
    S = s.courant
    lib().oracle_step(rows, cols, _ptr(f.h), _ptr(f.ex), _ptr(f.ey), _ptr(hn), _ptr(s.mu_coef), _ptr(s.eps_coef),
                      S, 1 - S, C.addressof(walls), len(live), C.addressof(srcs), last, len(rect),
                      rect.ctypes.data if len(rect) else None)
    f.h = hn
    f.step = n + 1
    ex, ey = f.ex, f.ey
    for i, j, series in f.probes:                                       # :246-255
        series.append(0.5 * ((s.eps[i][j] * (ex[i][j] ** 2 + ey[i][j] ** 2) + s.mu[i][j] * hn[i][j] ** 2)))
    for i, j, series in f.raw_probes:
        series.append((hn[i][j], ex[i][j], ey[i][j]))
    return hn


def run(f, s, times):
    for t in times:
        leapfrog(f, s, t)
    return f


def run_saving(f, s, times, step_frequency=5):
    saved = []
    for t in times:
        leapfrog(f, s, t)
        if f.step % step_frequency == 0:
            saved.append(f.h)
    return saved


def set_threads(n):
    """Best effort: cap the OpenMP team (no-op for a build without -fopenmp)."""
    try:
        C.CDLL("libgomp.so.1").omp_set_num_threads(int(n))
    except OSError:
        pass
