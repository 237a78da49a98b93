"""The streaming kernels' DEVICE CODE (fdtd_em_solver_b200/csrc/kernels.cuh: step_stream with its mixed/plain tile
paths, stepK_stream with its K-stage register pipeline, sacrificial halo lanes and prefetch ring) executed on the CPU
by a small SIMT emulator (tests/host_emul/simt.h: one warp at a time, every lane a fiber, shuffles through a slot
array) and compared bitwise with the oracle.

Both kernels are GPU-validated (tests/test_gpu_parity.py); what this adds is a way to pin an EDIT of them before it
costs GPU time -- together with tests/test_sass_manifest.py, which tells whether an edit changed the shipped code."""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

import stream_emul as E

pytestmark = pytest.mark.skipif(not E.available(), reason="needs g++ and the CUDA headers")

WALLS = list(itertools.product([False, True], repeat=4))


def _check(got, f, where, only=None):
    for name, a, b in zip(("h", "ex", "ey"), got, (f.h, f.ex, f.ey)):
        if only is not None:
            a, b = a[only[name]], b[only[name]]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s differs %s: %d cells, first %s" % (name, where, len(bad), bad[0]))


@pytest.mark.parametrize("vec,tile_rows,warps", [(2, 4, 1), (4, 5, 2), (2, 16, 4)])
@pytest.mark.parametrize("shape", [(3, 3), (9, 5), (33, 33), (40, 70), (67, 131)])
def test_step_stream_every_call_bitwise(shape, vec, tile_rows, warps):
    """All pulse kinds incl. the TF/SF line at column 0, reflectors, every wall mix: mixed and plain tiles."""
    rows, cols = shape
    for seed in (0, 1, 2):
        walls = WALLS[(seed * 5 + rows + cols) % 16]
        f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=4)
        for n, t in enumerate(times):
            got = E.step_stream(f, setup, times, n, vec=vec, tile_rows=tile_rows, warps=warps)
            R.leapfrog(f, setup, t)
            _check(got, f, "shape %s seed %d call %d" % (shape, seed, n))


def _plan(rows, cols, K, tile_rows, warps):
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, 0, rows
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, warps
    plan = L.FramePlan()
    assert lib.fdtd_plan_frame(C.byref(cfg), K, 0, None, C.byref(plan), None, None) == 0
    skip = np.zeros((max(plan.k_tiles_y, 1), plan.k_tiles_x), dtype=np.uint8)
    assert lib.fdtd_plan_frame(C.byref(cfg), K, 0, None, C.byref(plan), skip.ctypes.data, None) == 0
    return plan, skip


@pytest.mark.parametrize("variant", [0, 1, 2, 3], ids=["stepK_stream", "stepK_lean", "stepK_lean_umu", "stepK_skew_umu"])
@pytest.mark.parametrize("K,tile_rows", [(2, 8), (3, 12), (4, 16), (5, 7), (6, 16), (7, 9), (8, 24)])
def test_stepK_plain_tiles_bitwise(K, tile_rows, variant):
    """The K-stage pipeline on every tile the frame plan leaves to it: K calls in one pass == K oracle calls.  Both the
    shipped kernel and stepK_lean (row loop unrolled by two, running pointers, wrapping ring slots; odd tile heights
    exercise the remainder of the unrolled loop)."""
    rows, cols = 100, 230
    f, setup, times = SC.random_case(rows, cols, 9, n_steps=K, with_sources=False, with_rects=False)
    if variant >= 2:              # stepK_lean_umu / stepK_skew_umu: the mu_coef plane is ONE value, taken from the parameters
        setup.mu_coef = np.full_like(setup.mu_coef, setup.mu_coef[3, 5])
    plan, skip = _plan(rows, cols, K, tile_rows, 2)
    assert plan.k_tiles_y > 0 and not skip.all()
    got = E.stepK_stream(f, setup, times, 0, K, tile_rows, 2, plan.k_row_lo, plan.k_row_hi, skip, variant=variant)
    R.run(f, setup, times)
    covered = {}
    for name, arr in zip(("h", "ex", "ey"), (f.h, f.ex, f.ey)):
        m = np.zeros(arr.shape, dtype=bool)
        for ty, tx in np.argwhere(skip == 0):
            ta = plan.k_row_lo + ty * plan.k_tile_rows
            tb = min(ta + plan.k_tile_rows, plan.k_row_hi)
            m[ta:tb, tx * plan.k_tile_cols:(tx + 1) * plan.k_tile_cols] = True
        covered[name] = m
    assert covered["h"].any()
    _check(got, f, "K %d" % K, only=covered)
    for name, a in zip(("h", "ex", "ey"), got):
        assert np.isnan(a[~covered[name]]).all()              # nothing outside its tiles was stored
