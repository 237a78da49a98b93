"""GPU parity of stepK_lean_umu (fdtd_em_solver_b200/csrc/kernels.cuh): the plain K-step kernel for a UNIFORM mu_coef
plane -- mu_r = 1 everywhere, as in every script of the reference (/root/reference/tutorial.py, doubleslit.py, lens.py,
waveguide.py pass mu_r_max = 1 and never paint mu) and in the C5 recipe.  The library finds the property with a device
scan whenever the coefficients arrive (fdtd_upload_coeffs / fdtd_fill_synthetic_coeffs / fdtd_paint_coeffs /
fdtd_run_streamed; fdtd_rescan_coeffs after a direct write) and the kernel then takes the value from its parameters
instead of streaming the plane: 7 words per cell and pass instead of 8, the same bits.

Pinned on the CPU through the SIMT emulator (tests/test_stream_emulation.py, variant 2)."""
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = [pytest.mark.gpu]

WALLS = list(itertools.product([False, True], repeat=4))


def _engine_for(f, setup, times, dtype="float64", **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    S = setup.courant
    e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, dtype=dtype, **kw)
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


def _uniform(setup):
    setup.mu_coef = np.full_like(setup.mu_coef, setup.mu_coef[2, 3])
    return setup


@pytest.mark.parametrize("skew", [0, 1], ids=["stepK_lean_umu", "stepK_skew_umu"])
@pytest.mark.parametrize("temporal", [2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("shape,tile_rows,warps", [((200, 900), 8, 4), ((131, 517), 5, 2), ((300, 1500), 0, 0)])
def test_uniform_mu_groups_bitwise_against_the_oracle(shape, tile_rows, warps, temporal, skew, monkeypatch):
    """skew: FDTD_SKEW picks between the two kernels for a uniform plane -- the lean pipeline, and the one whose stage chain
    is cut at two boundaries (kernels.cuh stepK_skew_umu); the library's default is whichever the B200 measured faster."""
    monkeypatch.setenv("FDTD_SKEW", str(skew))
    rows, cols = shape
    for seed, k in ((1, 0), (2, 3), (3, 5)):
        f, setup, times = SC.random_case(rows, cols, seed, walls=WALLS[(seed * 3) % 16], n_steps=17)
        _uniform(setup)
        e = _engine_for(f, setup, times, kernel="stream", tile_rows=tile_rows, block_warps=warps, temporal=temporal)
        assert e.uniform_mu()
        if k:
            want = np.array(R.run_saving(f, setup, times, k))
            got = e.run(0, len(times), snapshot_every=k, total_calls=len(times))
            assert got.shape == want.shape and np.array_equal(got, want), (seed, k)
        else:
            R.run(f, setup, times)
            e.run(0, len(times))
        _same(e, f, "uniform mu, K %d, seed %d" % (temporal, seed))
        e.close()


def test_the_scan_sees_every_stored_cell_and_follows_the_planes(monkeypatch):
    """One differing cell anywhere (first row, last row, last column) keeps the general kernel; a direct write to the plane
    followed by fdtd_rescan_coeffs is honoured; FDTD_UNIFORM_MU=0 switches the fast path off; all with the oracle's bits."""
    rows, cols = 160, 700
    for cell in ((0, 0), (rows - 1, cols - 1), (rows // 2, 1), (0, cols - 1)):
        f, setup, times = SC.random_case(rows, cols, 5, walls=WALLS[6], n_steps=16)
        _uniform(setup)
        setup.mu_coef[cell] *= 1.0000001
        e = _engine_for(f, setup, times, temporal=8)
        assert not e.uniform_mu(), cell
        R.run(f, setup, times)
        e.run(0, len(times))
        _same(e, f, "one odd cell at %s" % (cell,))
        e.close()
    # direct writes to the caller-owned plane + rescan, both ways
    f, setup, times = SC.random_case(rows, cols, 6, walls=WALLS[9], n_steps=16)
    _uniform(setup)
    e = _engine_for(f, setup, times, temporal=8)
    assert e.uniform_mu()
    odd = setup.mu_coef.copy()
    odd[40:90, 100:400] *= 0.75
    import torch
    e.buffers[7][:rows, :cols].copy_(torch.from_numpy(odd))
    e.rescan_coefficients()
    assert not e.uniform_mu()
    setup.mu_coef = odd
    R.run(f, setup, times)
    e.run(0, len(times))
    _same(e, f, "plane rewritten to a non-uniform one")
    e.close()
    # the switch
    f, setup, times = SC.random_case(rows, cols, 7, walls=WALLS[3], n_steps=16)
    _uniform(setup)
    monkeypatch.setenv("FDTD_UNIFORM_MU", "0")
    e = _engine_for(f, setup, times, temporal=8)
    assert not e.uniform_mu()
    e.run(0, len(times))
    monkeypatch.delenv("FDTD_UNIFORM_MU")
    e2 = _engine_for(f, setup, times, temporal=8)
    assert e2.uniform_mu()
    e2.run(0, len(times))
    for a, b in zip(e.get_fields(), e2.get_fields()):
        assert np.array_equal(a, b)
    R.run(f, setup, times)
    _same(e2, f, "fast path vs general kernel vs oracle")
    e.close(); e2.close()


def test_float32_and_the_device_side_coefficient_paths():
    """float32 bits do not depend on the path; synthetic (C5) coefficients have mu_r = 1: the fast path is what bench.py runs."""
    from fdtd_em_solver_b200 import synthetic as SY
    f, setup, times = SC.random_case(150, 1200, 10, n_steps=16)
    _uniform(setup)
    outs = []
    for temporal in (1, 8):
        e = _engine_for(f, setup, times, dtype="float32", temporal=temporal)
        assert e.uniform_mu()
        e.run(0, len(times))
        outs.append(e.get_fields())
        e.close()
    for a, b in zip(*outs):
        assert np.array_equal(a, b)
    e, _ = SY.make_engine(1024, 2048, dtype="float64", n_calls=64, temporal=8)
    assert e.uniform_mu()
    e.run(0, 24)
    h8 = e.get_fields()[0]
    e.close()
    e, _ = SY.make_engine(1024, 2048, dtype="float64", n_calls=64, temporal=1)
    e.run(0, 24)
    assert np.array_equal(h8, e.get_fields()[0])
    e.close()


def test_streamed_runs_learn_the_property_at_the_end():
    """fdtd_run_streamed uploads the coefficient rows while the first groups already run: it steps with the general kernel
    and scans afterwards, so the NEXT run takes the fast path."""
    import ctypes as C
    from fdtd_em_solver_b200 import _lib as L
    rows, cols = 2048, 1500
    f, setup, times = SC.random_case(rows, cols, 4, walls=WALLS[0], n_steps=24)
    _uniform(setup)
    e = _engine_for(f, setup, times, temporal=8)
    eps = np.ascontiguousarray(setup.eps_coef); mu = np.ascontiguousarray(setup.mu_coef)
    out = np.empty((rows, cols))
    L.check(e.lib, e.handle, e.lib.fdtd_run_streamed(e.handle, eps.ctypes.data_as(C.c_void_p), mu.ctypes.data_as(C.c_void_p),
                                                     cols, 0, 0, 16, out.ctypes.data_as(C.c_void_p), cols))
    assert e.uniform_mu()
    e.run(16, 8)
    R.run(f, setup, times)
    _same(e, f, "streamed 16 calls + 8 calls on the fast path")
    e.close()


@pytest.mark.parametrize("skew", [0, 1])
def test_c5_slab_sized_columns_and_tall_tiles(skew, monkeypatch):
    """The geometry bench.py runs (512-row plain tiles, 32-row edge tiles, K = 8) on a grid tall enough for several tile
    rows, against single stepping on the device."""
    from fdtd_em_solver_b200 import synthetic as SY
    monkeypatch.setenv("FDTD_SKEW", str(skew))
    e8, _ = SY.make_engine(2048, 4096, dtype="float64", n_calls=128, temporal=8, tile_rows=512)
    e1, _ = SY.make_engine(2048, 4096, dtype="float64", n_calls=128, temporal=1)
    assert e8.uniform_mu() and e8.tuning()["tile_rows"] == 512
    e8.run(0, 43)
    e1.run(0, 43)
    for a, b in zip(e8.get_fields(), e1.get_fields()):
        assert np.array_equal(a, b)
    e8.close(); e1.close()
