"""Realtime view on the GPU (the reference loop is plot_and_show, solver.py:309-329): batches of calls
queued without waiting (fdtd_run_enqueue), decimated pictures of h fetched through the two-slot view ring
(fdtd_view_begin / fdtd_view_end) while the next batch runs.  Every picture must be the oracle's h after that batch, bit
for bit -- the pipeline changes when the host looks, never what it sees."""
import contextlib
import io
import sys
import types

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = pytest.mark.gpu


def _engine(f, setup, times, **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    e = Engine(rows, cols, courant=setup.courant, one_minus_courant=1 - setup.courant, reflect=setup.reflect_walls, **kw)
    e.upload_coefficients(setup.eps_coef, setup.mu_coef)
    e.set_sources(setup.pulses, times)
    e.set_reflectors(setup.reflectors)
    e.set_fields(f.h, f.ex, f.ey)
    return e


@pytest.mark.parametrize("kw,shape", [(dict(), (120, 200)),                                     # resident mode
                                      (dict(temporal=4, tile_rows=16, resident=False), (300, 700)),   # K-step groups + edge tiles
                                      (dict(tile_rows=8, resident=False), (90, 333))])           # single steps
@pytest.mark.parametrize("batch,strides", [(7, (1, 1)), (5, (3, 4))])
def test_pipelined_views_are_the_oracle_fields(kw, shape, batch, strides):
    rows, cols = shape
    n_steps = 4 * batch + 3
    f, setup, times = SC.random_case(rows, cols, 7, walls=(False, True, False, False), n_steps=n_steps)
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    e = _engine(f, setup, times, **kw)
    rs, cs = strides
    want, n = [], 0
    batches = []
    while n < n_steps:
        g = min(batch, n_steps - n)
        for t in times[n:n + g]:
            R.leapfrog(f, setup, t)
        want.append(f.h[::rs, ::cs].copy())
        batches.append((n, g))
        n += g
    got = []
    e.run_enqueue(*batches[0]); e.view_begin(rs, cs)
    for k in range(len(batches)):
        if k + 1 < len(batches):                       # the device runs one batch ahead of the picture being fetched
            e.run_enqueue(*batches[k + 1]); e.view_begin(rs, cs)
        got.append(e.view_end())
    e.sync()
    assert len(got) == len(want)
    for k, (a, b) in enumerate(zip(got, want)):
        assert a.shape == b.shape and np.array_equal(a, b), (k, np.argwhere(a != b)[:3])
    h, ex, ey = e.get_fields()
    assert np.array_equal(h, f.h) and np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
    assert e.stats()["steps"] == n_steps
    assert e.lib.fdtd_view_end(e.handle, None, 0, None, None) != 0          # nothing in flight any more: an error, no hang
    e.close()


def test_views_of_slabs_tile_the_global_picture():
    """Each slab pictures the rows of the GLOBAL decimated grid that it owns: stacked, they are h[::rs, ::cs]."""
    from fdtd_em_solver_b200 import slab as SL
    rows, cols, rs, cs = 200, 180, 3, 2
    f, setup, times = SC.random_case(rows, cols, 3, n_steps=1, with_sources=False, with_rects=False)
    parts = []
    for b, e_ in SL.split_rows(rows, 3):
        e = _engine(f, setup, times, row_begin=b, row_end=e_, resident=False)
        e.view_begin(rs, cs)
        parts.append(e.view_end())
        e.close()
    assert np.array_equal(np.concatenate(parts), f.h[::rs, ::cs])


def test_solver_realtime_path_with_frame_stride_and_decimation(monkeypatch):
    """The drop-in Solver's plot_and_show with frame_stride = 7, view_decimation = 2 (matplotlib replaced by a stand-in
    that plays every frame): pictures, final fields, step counter and recorded energies against the oracle."""
    shown = []

    class Image:
        def set_array(self, a): shown.append(np.array(a))

    class Axes:
        def imshow(self, *a, **k): return Image()
        def set_aspect(self, *a, **k): pass
        xaxis = types.SimpleNamespace(set_ticks_position=lambda *a: None, set_label_position=lambda *a: None)

    class Figure:
        def colorbar(self, *a, **k): pass
        def suptitle(self, *a, **k): pass

    plt = types.ModuleType("matplotlib.pyplot")
    plt.subplots = lambda *a, **k: (Figure(), Axes())
    plt.show = lambda *a, **k: None
    anim = types.ModuleType("matplotlib.animation")

    def FuncAnimation(fig, func, frames, **k):
        for frame in frames:
            func(frame)
    anim.FuncAnimation = FuncAnimation
    top = types.ModuleType("matplotlib")
    top.pyplot, top.animation = plt, anim
    for name, mod in (("matplotlib", top), ("matplotlib.pyplot", plt), ("matplotlib.animation", anim)):
        monkeypatch.setitem(sys.modules, name, mod)
    from fdtd_em_solver_b200 import solver as S

    sc = SC.spec("waveguide")
    f, setup, times, _ = SC.restate(sc)
    with contextlib.redirect_stdout(io.StringIO()):
        s = SC.drive(S, sc)
        s.end_time = 200.5 * s.dt
        s.frame_stride, s.view_decimation = 7, 2
        s.solve()
    want, energies = [], []
    i, j = sc["probe_cells"][0]
    for n, t in enumerate(times[:201]):
        R.leapfrog(f, setup, t)
        if (n + 1) % 7 == 0 or n == 200:
            want.append(f.h[::2, ::2].copy())
    assert s.step == 201 and len(shown) == len(want) == 29
    assert all(np.array_equal(a, b) for a, b in zip(shown, want))
    assert np.array_equal(s.h, f.h) and np.array_equal(s.ex, f.ex)
    if s.recorded_energies:
        assert len(s.recorded_energies[0]["energies"]) == 201
