"""TEST INFRASTRUCTURE (see oracle/__init__.py): the reference step (oracle/restated.py, /root/reference/solver.py:169-257)
with the fields passed through SPLIT STORAGE after every call -- what the library's dtype FDTD_F32X ("float32x") computes:
float64 coefficients and arithmetic in the reference's order, every stored field value rounded to
    hi + lo,   hi = float32(x),   lo = bfloat16(float32(x - hi)),   both round-to-nearest-even
(include/fdtd2d.h; fdtd_em_solver_b200/csrc/kernels.cuh split_encode / split_decode).  The reference has no such mode:
this restates the library's storage format on top of the reference's arithmetic, so that the CUDA path can be held to
it bit for bit, and measures its distance from the float64 reference (the 1e-5 gate of BASELINE.json's north_star)."""
import numpy as np

from . import restated as R


def bf16_round(r32):
    """float32 array -> the nearest bfloat16 (ties to even), returned as float32."""
    u = np.ascontiguousarray(r32, dtype=np.float32).view(np.uint32).astype(np.uint64)
    finite = (u & 0x7F800000) != 0x7F800000
    rounded = np.where(finite, u + 0x7FFF + ((u >> 16) & 1), u) >> 16 << 16
    return rounded.astype(np.uint32).view(np.float32)


def split_round(x):
    """A float64 array as it comes back from split storage."""
    x = np.asarray(x, dtype=np.float64)
    hi = x.astype(np.float32)
    lo = bf16_round((x - hi.astype(np.float64)).astype(np.float32))
    return hi.astype(np.float64) + lo.astype(np.float64)


def store(f):
    f.h, f.ex, f.ey = split_round(f.h), split_round(f.ex), split_round(f.ey)


def leapfrog(f, s, time):
    """One call: the reference's update on the stored fields, results back through storage."""
    R.leapfrog(f, s, time)
    store(f)
    # record_energy (solver.py:246-255) reads the arrays the call LEFT: in this mode those are the stored values
    for i, j, series in f.probes:
        series[-1] = 0.5 * ((s.eps[i][j] * (f.ex[i][j] ** 2 + f.ey[i][j] ** 2) + s.mu[i][j] * f.h[i][j] ** 2))
    for i, j, series in f.raw_probes:
        series[-1] = (f.h[i][j], f.ex[i][j], f.ey[i][j])


def run(f, s, times):
    store(f)                                   # the initial fields are uploaded into split storage
    for t in times:
        leapfrog(f, s, t)


def run_saving(f, s, times, step_frequency=5):
    """plot_and_save (solver.py:331-337) incl. the list-aliasing rule (App. A-Q8): the saved array is the one the next
    call's hard point-source write lands in -- a float64 value written into a float64 snapshot, as in the reference."""
    store(f)
    saved = []
    for t in times:
        leapfrog(f, s, t)
        if f.step % step_frequency == 0:
            saved.append(f.h)
    return saved
