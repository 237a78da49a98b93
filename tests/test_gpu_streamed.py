"""fdtd_run_streamed on the GPU: upload -> calls -> download pipelined over row bands must give the bits of the four
separate calls (and therefore of the oracle), in the single-sweep form (short runs) and in the three-phase form (long
runs: diagonal sweep, whole-grid groups, diagonal sweep), with every feature on the grid."""
import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = pytest.mark.gpu


def _engine(f, setup, times, **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    e = Engine(rows, cols, courant=setup.courant, one_minus_courant=1 - setup.courant, reflect=setup.reflect_walls,
               resident=False, **kw)
    e.set_sources(setup.pulses, times)
    e.set_reflectors(setup.reflectors)
    return e


@pytest.mark.parametrize("shape,tile_rows,K,n_steps", [
    ((300, 700), 16, 8, 20),        # 10 bands, 3 groups (8 + 8 + 4): one sweep
    ((300, 700), 16, 4, 19),        # 4+4+4+4+3
    ((100, 333), 16, 2, 21),        # 4 bands, 10 groups (2 x 9 + 3 -> ... ) : three phases
    ((130, 400), 8, 3, 40),         # bands of 24 rows, 14 groups: three phases
    ((1000, 900), 128, 8, 17),      # 8 + 7 + 2 (no single step)
    ((97, 150), 4, 5, 33)])
def test_streamed_run_is_the_sequential_run(shape, tile_rows, K, n_steps, monkeypatch):
    monkeypatch.setenv("FDTD_STREAMED_STRICT", "1")                  # falling back to the four separate calls is an error here
    rows, cols = shape
    f, setup, times = SC.random_case(rows, cols, 13, walls=(False, True, False, False), n_steps=n_steps)
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    f.h[:] = 0; f.ex[:] = 0; f.ey[:] = 0                             # the streamed run starts from zero fields
    cells = [(0, 0), (rows // 2, cols // 2), (rows - 1, cols - 1)]
    e = _engine(f, setup, times, tile_rows=tile_rows, temporal=K)
    e.set_probes(cells)
    h = e.run_streamed(setup.eps_coef, setup.mu_coef, 0, n_steps)
    want_rec = np.zeros((n_steps, len(cells), 3))
    for n, t in enumerate(times):
        R.leapfrog(f, setup, t)
        for k, (i, j) in enumerate(cells):
            want_rec[n, k] = (f.h[i, j], f.ex[i, j], f.ey[i, j])
    assert np.array_equal(h, f.h), np.argwhere(h != f.h)[:4]
    h2, ex2, ey2 = e.get_fields()
    assert np.array_equal(h2, f.h) and np.array_equal(ex2, f.ex) and np.array_equal(ey2, f.ey)
    assert np.array_equal(e.probes(), want_rec)
    assert e.stats()["steps"] == n_steps
    e.close()


def test_streamed_continuation_and_pinned_buffers(monkeypatch):
    """zero_fields = False continues from the device state; pinned torch buffers take the asynchronous route."""
    import torch
    monkeypatch.setenv("FDTD_STREAMED_STRICT", "1")
    rows, cols, K = 400, 520, 8
    f, setup, times = SC.random_case(rows, cols, 4, n_steps=40)
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    e = _engine(f, setup, times, tile_rows=16, temporal=K)
    e.upload_coefficients(setup.eps_coef, setup.mu_coef)
    e.set_fields(f.h, f.ex, f.ey)
    e.run(0, 11)
    eps = torch.from_numpy(setup.eps_coef.copy()).pin_memory()
    mu = torch.from_numpy(setup.mu_coef.copy()).pin_memory()
    out = torch.zeros((rows, cols), dtype=torch.float64).pin_memory()
    h = e.run_streamed(eps.numpy(), mu.numpy(), 11, 29, zero_fields=False, h_out=out.numpy())
    R.run(f, setup, times)
    assert np.array_equal(h, f.h)
    _, ex, ey = e.get_fields()
    assert np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
    e.close()


def test_streamed_falls_back_where_the_pipeline_does_not_apply(monkeypatch):
    """float32, resident-sized defaults, a wrapping TF/SF line: the same call runs the four steps one after the other."""
    rows, cols = 120, 200
    f, setup, times = SC.random_case(rows, cols, 3, n_steps=12)          # seed 3 keeps a "right" line at column 0
    f.h[:] = 0; f.ex[:] = 0; f.ey[:] = 0
    e = _engine(f, setup, times, tile_rows=16, temporal=4)
    assert any(p.direction == "right" and p.col == 0 for p in setup.pulses)
    h = e.run_streamed(setup.eps_coef, setup.mu_coef, 0, 12)
    R.run(f, setup, times)
    assert np.array_equal(h, f.h)
    e.close()
    monkeypatch.setenv("FDTD_STREAMED_STRICT", "1")
    from fdtd_em_solver_b200._lib import FdtdError
    e = _engine(f, setup, times, tile_rows=16, temporal=4)
    with pytest.raises(FdtdError):
        e.run_streamed(setup.eps_coef, setup.mu_coef, 0, 12)
    e.close()
