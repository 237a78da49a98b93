"""CPU check of step_resident's window algorithm (fdtd_em_solver_b200/csrc/resident.cuh) against the oracle.

The kernel's per-cell phase functions are __host__ __device__; tests/host_emul/resident_emul.cu runs them with plain
loops, so the tile / apron / phase logic is pinned bit-for-bit here, without a GPU:

  * every geometry incl. 1x1 tiles (every stored value then comes through the 2-cell apron),
  * three cell orders inside a phase (the result must not depend on thread order),
  * all 16 wall combinations, all pulse kinds, clipped / overlapping reflectors, probes, rectangular grids,
  * the four BASELINE.json script scenes in launches of `step_frequency` calls, as fdtd_run issues them.

The -m gpu tests (tests/test_gpu_resident.py) cover what an emulation cannot: the cooperative launch and grid.sync.
"""
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

import resident_emul as E

pytestmark = pytest.mark.skipif(not E.available(), reason="nvcc not found (the emulation is compiled from resident.cuh)")

WALLS = list(itertools.product([False, True], repeat=4))


def _eligible(setup):
    """fdtd2d.cu resident_eligible: a TF/SF "right" line at column 0 stays on the per-call path."""
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    return setup


def _check(got, f, where):
    for name, a, b in zip(("h", "ex", "ey"), got[:3], (f.h, f.ex, f.ey)):
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s differs %s: %d cells, first %s: got %r want %r" % (
                name, where, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


@pytest.mark.parametrize("geom", [0, 1, 2, 3])
@pytest.mark.parametrize("shape", [(3, 3), (4, 7), (9, 5), (33, 33), (40, 70), (67, 131)])
def test_random_cases_bitwise_every_geometry(geom, shape):
    rows, cols = shape
    for seed in range(4):
        walls = WALLS[(seed * 5 + rows + cols) % 16]
        f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=10)
        setup = _eligible(setup)
        cells = [(0, 0), (rows - 1, cols - 1), (rows // 2, cols // 3)]
        f.raw_probes = [(i, j, []) for i, j in cells]
        got = E.emulate(f, setup, times, 0, len(times), geom=geom, order=seed % 3, probe_cells=cells)
        for t in times:
            R.leapfrog(f, setup, t)
        _check(got, f, "shape %s seed %d walls %s geom %d" % (shape, seed, walls, geom))
        for c, (i, j, series) in enumerate(f.raw_probes):
            assert np.array_equal(got[3][:, c, :], np.array(series))


@pytest.mark.parametrize("walls", WALLS)
def test_all_wall_combinations_each_call(walls):
    """One call per launch, chained: every intermediate state is compared."""
    for shape, geom in (((6, 6), 1), ((17, 12), 2), ((12, 70), 0)):
        f, setup, times = SC.random_case(shape[0], shape[1], 11, walls=walls, n_steps=6)
        setup = _eligible(setup)
        for n, t in enumerate(times):
            got = E.emulate(f, setup, times, n, 1, geom=geom, order=n % 3)
            R.leapfrog(f, setup, t)
            _check(got, f, "walls %s shape %s call %d" % (walls, shape, n))


def test_phase_cell_order_does_not_matter():
    f, setup, times = SC.random_case(37, 90, 2, walls=(False, True, False, False), n_steps=9)
    setup = _eligible(setup)
    runs = [E.emulate(f, setup, times, 0, len(times), geom=0, order=o) for o in (0, 1, 2)]
    for other in runs[1:]:
        for a, b in zip(runs[0][:3], other[:3]):
            assert np.array_equal(a, b)


@pytest.mark.parametrize("name,n_calls", [("tutorial", 400), ("lens", 400), ("waveguide", 600),
                                          ("waveguide_mixed_walls", 400), ("mixed", 400), ("doubleslit", 300)])
def test_script_scenes_in_launches_of_step_frequency(name, n_calls):
    """The kernel geometry on the reference's own scenes, advanced the way fdtd_run does in resident mode: one launch
    per `step_frequency` calls, the result of a launch feeding the next."""
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    k = sc["step_frequency"]
    cells = [(i, j) for i, j, _ in f.probes]
    f.raw_probes = [(i, j, []) for i, j in cells]
    g = R.Fields(f.h.copy(), f.ex.copy(), f.ey.copy())
    records = []
    for n in range(0, n_calls, k):
        h, ex, ey, rec = E.emulate(g, setup, times, n, k, geom=0, order=(n // k) % 3, probe_cells=cells)
        g = R.Fields(h, ex, ey)
        records.append(rec)
        for t in times[n:n + k]:
            R.leapfrog(f, setup, t)
        if (n // k) % 20 == 0 or n + k >= n_calls:
            _check((h, ex, ey), f, "%s after call %d" % (name, n + k))
    assert np.abs(f.h).max() > 0
    rec = np.concatenate(records)
    for c, (i, j, series) in enumerate(f.raw_probes):
        assert np.array_equal(rec[:, c, :], np.array(series))
