"""GPU parity of stepK_lean (fdtd_em_solver_b200/csrc/kernels.cuh): the K-step kernel with the row loop unrolled by two
and running pointers (about 30 % fewer issue slots per row), selected with fdtd_set_kstep_variant(h, 1).  Same operations
in the same order as stepK_stream, so everything is bit-exact against the oracle / against single stepping.

Pinned on the CPU through the SIMT emulator (tests/test_stream_emulation.py).  No GPU run yet when round 1's GPU budget
ended, so these only run with FDTD_TEST_KSTEP_LEAN=1 and stepK_stream stays the default.
"""
import itertools
import os

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = [pytest.mark.gpu]           # first B200 run: round 2, profiles/r2_first_run_new_kernels_tests.log (all green)

WALLS = list(itertools.product([False, True], repeat=4))


def _engine_for(f, setup, times, **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    S = setup.courant
    e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, **kw)
    e.upload_coefficients(setup.eps_coef, setup.mu_coef)
    e.set_sources(setup.pulses, times)
    e.set_reflectors(setup.reflectors)
    e.set_fields(f.h, f.ex, f.ey)
    return e


def _same(e, f, where=""):
    h, ex, ey = e.get_fields()
    for name, a, b in (("h", h, f.h), ("ex", ex, f.ex), ("ey", ey, f.ey)):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s differs %s: %d cells, first %s: got %r want %r" % (
                name, where, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


@pytest.mark.parametrize("temporal", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("shape,tile_rows,warps", [((200, 900), 8, 4), ((131, 517), 5, 2), ((64, 300), 16, 1)])
def test_k_steps_per_pass_bitwise(shape, tile_rows, warps, temporal):
    rows, cols = shape
    for seed, k in ((1, 0), (2, 3), (3, 5)):
        f, setup, times = SC.random_case(rows, cols, seed, walls=WALLS[(seed * 3) % 16], n_steps=17)
        e = _engine_for(f, setup, times, kernel="stream", tile_rows=tile_rows, block_warps=warps, temporal=temporal,
                        kstep_variant=1)
        if k:
            want = np.array(R.run_saving(f, setup, times, k))
            got = e.run(0, len(times), snapshot_every=k, total_calls=len(times))
            assert got.shape == want.shape and np.array_equal(got, want), (seed, k)
        else:
            R.run(f, setup, times)
            e.run(0, len(times))
        _same(e, f, "stepK_lean, K %d, seed %d" % (temporal, seed))
        e.close()


def test_interior_only_and_fp32():
    from fdtd_em_solver_b200.engine import Engine
    for temporal in (2, 3, 4, 5, 6, 7, 8):
        f, setup, times = SC.random_case(300, 2000, 9, n_steps=19, with_sources=False, with_rects=False)
        e = _engine_for(f, setup, times, kernel="stream", tile_rows=8, temporal=temporal, kstep_variant=1)
        R.run(f, setup, times)
        e.run(0, len(times))
        _same(e, f, "interior %d-step" % temporal)
        e.close()
    f, setup, times = SC.random_case(150, 1200, 10, n_steps=10)
    outs = []
    for temporal, variant in ((1, 0), (2, 1), (4, 1), (6, 1)):
        rows, cols = f.h.shape
        S = setup.courant
        e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, dtype="float32",
                   tile_rows=8, temporal=temporal, kstep_variant=variant)
        e.upload_coefficients(setup.eps_coef, setup.mu_coef); e.set_sources(setup.pulses, times)
        e.set_reflectors(setup.reflectors); e.set_fields(f.h, f.ex, f.ey)
        e.run(0, len(times))
        outs.append(e.get_fields())
        e.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


def test_2048_grid_default_geometry_vs_oracle():
    f, setup, times = SC.random_case(2048, 2048, 21, walls=(False, False, False, False), n_steps=20)
    e = _engine_for(f, setup, times, kernel="stream", temporal=6, kstep_variant=1)
    R.run(f, setup, times)
    e.run(0, len(times))
    _same(e, f, "2048^2, K=6, stepK_lean")
    e.close()


def test_full_size_slab_equals_single_stepping():
    import torch
    from fdtd_em_solver_b200 import synthetic as SY
    free, total = torch.cuda.mem_get_info()
    if free < 110 * 2 ** 30:
        pytest.skip("needs ~100 GB of device memory")
    rows, cols, n = 8192, 65536, 19
    a, _ = SY.make_engine(rows, cols, n_calls=64, temporal=1)
    b, _ = SY.make_engine(rows, cols, n_calls=64, temporal=6, kstep_variant=1)
    a.run(0, n); b.run(0, n)
    for x, y in zip(a.generation_buffers(), b.generation_buffers()):
        assert torch.equal(x, y)
    assert float(a.generation_buffers()[0].abs().max()) > 0
    a.close(); b.close()
