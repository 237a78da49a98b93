"""GPU parity of frame mode 3 (fdtd_em_solver_b200/csrc/edge.cuh): every tile of a temporal-blocking group that the
plain K-step kernel leaves out -- wall bands and columns, pulse lines, sources, reflectors, probes -- advanced K calls
by ONE stepK_edge launch (the K-step register pipeline with those features as tile predicates) instead of K masked
single-step passes (mode 0, which these tests run beside it).  Everything is bit-exact against the oracle (fp64 bar of
BASELINE.json: 1e-12 relative L2; what is asserted is equality).

The kernel's device code is pinned on the CPU by tests/test_edge_emulation.py (SIMT emulator).  First B200 run: round 2,
profiles/r2_first_run_stepK_edge_tests.log.
"""
import itertools
import os

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = [pytest.mark.gpu]

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


def _without_wrapping_lines(setup):
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    return setup


@pytest.mark.parametrize("mode", [0, 3])
@pytest.mark.parametrize("temporal", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("shape,tile_rows,warps", [((200, 900), 8, 4), ((131, 517), 4, 2), ((64, 300), 16, 1)])
def test_k_steps_per_pass_both_frame_modes(shape, tile_rows, warps, temporal, mode):
    rows, cols = shape
    for seed, k in ((1, 0), (2, 3), (4, 5), (5, 2)):
        f, setup, times = SC.random_case(rows, cols, seed, walls=WALLS[(seed * 3) % 16], n_steps=17)
        setup = _without_wrapping_lines(setup)
        cells = [(0, 0), (rows // 2, cols // 2), (rows - 1, cols - 1), (5, cols // 3)]
        f.raw_probes = [(i, j, []) for i, j in cells]
        e = _engine_for(f, setup, times, kernel="stream", tile_rows=tile_rows, block_warps=warps, temporal=temporal,
                        frame_mode=mode)
        e.set_probes(cells)
        launches0 = e.stats()["kernel_launches"]
        if k:
            want = np.array(R.run_saving(f, setup, times, k))
            got = e.run(0, len(times), snapshot_every=k, total_calls=len(times))
            assert got.shape == want.shape and np.array_equal(got, want), (seed, k)
        else:
            R.run(f, setup, times)
            e.run(0, len(times))
            groups, rest = divmod(len(times), temporal)
            if mode == 3:
                # plain kernel + one stepK_edge launch per tile class present (<= 5): never a chain of K masked passes
                assert e.stats()["kernel_launches"] - launches0 <= 6 * (groups + 1) + 1     # (+ the remainder group / step)
        _same(e, f, "frame mode %d, K %d, seed %d" % (mode, temporal, seed))
        raw = e.probes()
        for c, (i, j, series) in enumerate(f.raw_probes):
            assert np.array_equal(raw[:, c, :], np.array(series)), (seed, c)
        e.close()


@pytest.mark.parametrize("mode", [0, 3])
def test_wrapping_line_keeps_the_masked_passes(mode):
    """A TF/SF "right" line at column 0 (python's h[:, col-1] wraps to the far column) is outside stepK_edge: the group
    silently uses the masked passes, or -- without scratch generations, the default of mode 3 -- single steps."""
    f, setup, times = SC.random_case(131, 517, 3, n_steps=13)          # seed % 3 == 0: has the column-0 line
    assert any(p.direction == "right" and p.col == 0 for p in setup.pulses)
    e = _engine_for(f, setup, times, kernel="stream", tile_rows=8, temporal=4, frame_mode=mode)
    R.run(f, setup, times)
    e.run(0, len(times))
    _same(e, f, "wrapping line")
    e.close()


@pytest.mark.parametrize("mode", [0, 3])
def test_fp32_same_bits_as_single_stepping(mode):
    from fdtd_em_solver_b200.engine import Engine
    f, setup, times = SC.random_case(150, 1200, 10, n_steps=13)
    setup = _without_wrapping_lines(setup)
    outs = []
    for temporal, fm in ((1, 3), (6, mode), (4, mode)):
        rows, cols = f.h.shape
        S = setup.courant
        e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, dtype="float32",
                   tile_rows=8, temporal=temporal, frame_mode=fm)
        e.upload_coefficients(setup.eps_coef, setup.mu_coef); e.set_sources(setup.pulses, times)
        e.set_reflectors(setup.reflectors); e.set_fields(f.h, f.ex, f.ey)
        e.run(0, len(times))
        outs.append(e.get_fields())
        e.close()
    for other in outs[1:]:
        for a, b in zip(outs[0], other):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("mode", [0, 3])
def test_2048_grid_default_geometry_vs_oracle(mode):
    f, setup, times = SC.random_case(2048, 2048, 21, walls=(False, False, False, False), n_steps=20)
    setup = _without_wrapping_lines(setup)
    e = _engine_for(f, setup, times, kernel="stream", temporal=6, frame_mode=mode)
    R.run(f, setup, times)
    e.run(0, len(times))
    _same(e, f, "2048^2, K=6, frame mode %d" % mode)
    e.close()


def test_full_size_slab_equals_single_stepping():
    """BASELINE.json's per-GPU workload (8192 x 65536 fp64, C5 recipe) with the library's defaults (K = 8, stepK_lean +
    stepK_edge, 512-row plain tiles) must leave the bits of plain single stepping (compared on the device)."""
    import torch
    from fdtd_em_solver_b200 import synthetic as SY
    free, total = torch.cuda.mem_get_info()
    if free < 70 * 2 ** 30:
        pytest.skip("needs ~70 GB of device memory")
    rows, cols, n = 8192, 65536, 21
    a, _ = SY.make_engine(rows, cols, n_calls=64, temporal=1)
    b, _ = SY.make_engine(rows, cols, n_calls=64, temporal=8)
    assert b.tuning()["frame_mode"] == 3 and b.tuning()["kstep_variant"] == 1 and len(b.buffers) == 8
    a.run(0, n); b.run(0, n)
    for x, y in zip(a.generation_buffers(), b.generation_buffers()):
        assert torch.equal(x, y)
    assert float(a.generation_buffers()[0].abs().max()) > 0
    a.close(); b.close()
