"""dtype "float32x" (FDTD_F32X) on the GPU: fields stored as fp32 hi + bf16 lo planes (6 bytes per value), float64
coefficients and arithmetic.  Two bars:

* BIT-EXACT against oracle/split_storage.py (the reference's step, oracle/restated.py, with the fields passed through the
  storage format after every call) -- every call, groups of calls, snapshots, probes;
* BASELINE.json's gate for the reduced-precision path: relative L2 <= 1e-5 against the float64 reference on EVERY snapshot
  of the reference's scripts, and on the probe series.  The float64 side is the library's float64 path, which the parity
  tests hold bit-identical to /root/reference/solver.py:169-257."""
import contextlib
import io
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC, split_storage as SS

pytestmark = pytest.mark.gpu

GATE = 1e-5                      # BASELINE.json north_star: "the optional fp32 path must agree within 1e-5 relative L2"
WALLS = list(itertools.product([False, True], repeat=4))


def _engine(f, setup, times, **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    e = Engine(rows, cols, courant=setup.courant, one_minus_courant=1 - setup.courant, reflect=setup.reflect_walls,
               dtype="float32x", **kw)
    e.upload_coefficients(setup.eps_coef, setup.mu_coef)
    e.set_sources(setup.pulses, times)
    e.set_reflectors(setup.reflectors)
    e.set_fields(f.h, f.ex, f.ey)
    return e


def _same(e, f, where):
    h, ex, ey = e.get_fields()
    for name, a, b in (("h", h, f.h), ("ex", ex, f.ex), ("ey", ey, f.ey)):
        assert np.array_equal(a, b), "%s differs %s: first %s" % (name, where, np.argwhere(a != b)[:3])


def _case(rows, cols, seed, n_steps, walls):
    f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=n_steps)
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    return f, setup, times


@pytest.mark.parametrize("shape", [(3, 3), (9, 5), (33, 33), (40, 70), (67, 131), (150, 400)])
def test_every_call_bitwise_against_the_split_oracle(shape):
    rows, cols = shape
    for seed in range(3):
        walls = WALLS[(seed * 5 + rows + cols) % 16]
        f, setup, times = _case(rows, cols, seed, 10, walls)
        e = _engine(f, setup, times)
        SS.store(f)
        _same(e, f, "after the upload")
        for n, t in enumerate(times):
            SS.leapfrog(f, setup, t)
            e.step(n)
            _same(e, f, "shape %s seed %d walls %s call %d" % (shape, seed, walls, n))
        e.close()


@pytest.mark.parametrize("snap", [0, 3, 5, 8])
def test_runs_snapshots_and_probes_bitwise_against_the_split_oracle(snap):
    rows, cols, n_steps = 131, 517, 29
    for seed, walls in ((1, WALLS[0]), (2, WALLS[9]), (4, WALLS[15])):
        f, setup, times = _case(rows, cols, seed, n_steps, walls)
        cells = [(0, 0), (rows // 2, cols // 2), (rows - 1, cols - 1), (5, cols // 3)]
        f.raw_probes = [(i, j, []) for i, j in cells]
        e = _engine(f, setup, times)
        e.set_probes(cells)
        if snap:
            got = e.run(0, n_steps, snapshot_every=snap, total_calls=n_steps)
            want = np.array(SS.run_saving(f, setup, times, snap))
            assert got.shape == want.shape and np.array_equal(got, want), (seed, snap)
        else:
            e.run(0, n_steps)
            SS.run(f, setup, times)
        _same(e, f, "run seed %d snapshots every %d" % (seed, snap))
        raw = e.probes()
        for c, (i, j, series) in enumerate(f.raw_probes):
            assert np.array_equal(raw[:, c, :], np.array(series)), (seed, c)
        if snap in (0, 8):                   # 29 calls = 3 groups of 8 + one of 5, each a plain-tile launch + the edge-tile
            assert e.stats()["kernel_launches"] < 2 * n_steps       # launches of its near / far halves (+ snapshot packs)
        e.close()


def test_views_and_a_4096_grid():
    """The realtime view decodes split storage; a grid beyond L2 with every feature stays on the oracle's bits."""
    rows, cols, n_steps = 1200, 4096, 11
    f, setup, times = _case(rows, cols, 7, n_steps, WALLS[5])
    e = _engine(f, setup, times)
    e.run_enqueue(0, n_steps)
    e.view_begin(3, 5)
    pic = e.view_end()
    SS.run(f, setup, times)
    assert np.array_equal(pic, f.h[::3, ::5])
    _same(e, f, "1200 x 4096")
    e.close()


def test_what_the_mode_does_not_take_is_refused_loudly():
    from fdtd_em_solver_b200.engine import Engine
    from fdtd_em_solver_b200._lib import FdtdError
    with pytest.raises(FdtdError):                               # y-slabs
        Engine(100, 100, courant=0.1, dtype="float32x", row_begin=0, row_end=50)
    f, setup, times = SC.random_case(60, 90, 3, n_steps=4)       # seed 3 keeps a TF/SF "right" line at column 0
    assert any(p.direction == "right" and p.col == 0 for p in setup.pulses)
    e = _engine(f, setup, times)
    with pytest.raises(FdtdError):
        e.run(0, 4)
    e.close()


def _stacks(name, dtype, n_calls=None):
    """Snapshot stack of a reference script through the drop-in Solver, and the energy-probe series at its probe cell
    (the on-device probes: every call, not only the snapshot cadence)."""
    from fdtd_em_solver_b200 import solver as S
    sc = SC.spec(name)
    with contextlib.redirect_stdout(io.StringIO()):
        s = SC.drive(S, sc, dtype=dtype)
        if not s.recorded_energies:
            i, j = sc["probe_cells"][0]
            s.recorded_energies.append({"i": i, "j": j, "location": None, "energies": []})
        s._coefficients()
        s.time_vector = np.arange(0, s.end_time, s.dt)
        if n_calls:
            s.time_vector = s.time_vector[:n_calls]
        stack = s._run_device(s.time_vector, sc["step_frequency"])
    return stack, np.array(s.recorded_energies[0]["energies"])


@pytest.mark.parametrize("name,n_calls", [("tutorial", None), ("lens", None), ("waveguide", None), ("doubleslit", 12000)])
def test_the_1e5_gate_on_every_snapshot_of_the_reference_scripts(name, n_calls):
    """C1, C3, C4 at full length and the first 12 000 calls of C2 (where plain float32 is at 3.3e-4)."""
    ref, ref_series = _stacks(name, "float64", n_calls)
    got, got_series = _stacks(name, "float32x", n_calls)
    assert got.shape == ref.shape and len(ref) > 100
    den = np.sqrt((ref ** 2).sum(axis=(1, 2)))
    num = np.sqrt(((ref - got) ** 2).sum(axis=(1, 2)))
    live = den > 0
    rel = np.zeros_like(den)
    rel[live] = num[live] / den[live]
    assert (num[~live] == 0).all()
    assert rel.max() <= GATE, "%s: worst snapshot %d at %.3g" % (name, int(rel.argmax()), rel.max())
    assert np.linalg.norm(got_series - ref_series) <= GATE * np.linalg.norm(ref_series)
    print("%s: %d snapshots, worst relative L2 %.3g, probe series %.3g" % (
        name, len(ref), rel.max(), np.linalg.norm(got_series - ref_series) / np.linalg.norm(ref_series)))
