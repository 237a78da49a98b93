"""CPU: the drop-in Solver's host logic end to end, with the device engine replaced by the oracle-backed test double
(tests/fake_engine.py).  What is checked here is solver.py itself: engine life-cycle, run loops, snapshot and probe
plumbing, update() on and off the time axis, reset / re-solve, and the reference's error behaviour."""
import contextlib
import io

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC
from fake_engine import FakeEngine


@pytest.fixture
def S(monkeypatch):
    from fdtd_em_solver_b200 import solver
    monkeypatch.setattr(solver, "Engine", FakeEngine)
    return solver


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


@pytest.mark.parametrize("name,n_calls", [("tutorial", 600), ("waveguide", 900), ("mixed", 700), ("lens", 400)])
def test_save_path_snapshots_probes_and_files(S, name, n_calls, tmp_path):
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    times = times[:n_calls]
    s = _quiet(SC.drive, S, sc)
    s.end_time = float(times[-1]) + 0.5 * s.dt
    s.save(str(tmp_path / name))
    _quiet(s.solve, realtime=False, step_frequency=sc["step_frequency"])
    want = np.array(R.run_saving(f, setup, times, sc["step_frequency"]))
    got = np.load(str(tmp_path / name) + ".npy")
    assert got.shape == want.shape and np.array_equal(got, want)
    assert len(s.h_list) == len(want) and np.array_equal(s.h_list[-1], want[-1])
    assert np.array_equal(s.h, f.h) and np.array_equal(s.ex, f.ex) and np.array_equal(s.ey, f.ey)
    assert s.step == len(times) and len(s.time_vector) == len(times)
    for d, (i, j, series) in zip(s.recorded_energies, f.probes):
        assert d["energies"] == series
    import json
    meta = json.load(open(str(tmp_path / name) + ".txt"))
    assert meta["dt"] == s.dt and meta["step_frequency"] == sc["step_frequency"] and len(meta["material"]) == s.size


def test_update_on_and_off_the_time_axis(S):
    sc = SC.spec("waveguide")
    f, setup, times, info = SC.restate(sc)
    s = _quiet(SC.drive, S, sc)
    for t in times[:40]:                                    # on-axis: exactly the loop of solver.py:289-291
        R.leapfrog(f, setup, t)
        assert np.array_equal(s.update(t), f.h)
    for k in range(5):                                      # off-axis: the caller's time gates the pulses, step indexes the tables
        t = float(times[40 + k]) * 1.0003 + 1e-14
        R.leapfrog(f, setup, t)
        assert np.array_equal(s.update(t), f.h)
    assert np.array_equal(s.ex, f.ex) and np.array_equal(s.ey, f.ey)
    assert s.recorded_energies[0]["energies"] == f.probes[0][2] and s.step == 45


def test_reset_keeps_history_and_rezeroes_fields(S):
    sc = SC.spec("waveguide")
    s = _quiet(SC.drive, S, sc)
    s.end_time = 300.5 * s.dt
    _quiet(s.solve, solve_only=True)
    first = s.h.copy()
    assert s.step == 301 and first.any() and len(s.recorded_energies[0]["energies"]) == 301
    s.reset()                                               # solver.py:347-349: step and fields only
    assert s.step == 0 and not s.h.any() and len(s.recorded_energies[0]["energies"]) == 301
    created = FakeEngine.created
    _quiet(s.solve, solve_only=True)
    assert np.array_equal(s.h, first) and len(s.recorded_energies[0]["energies"]) == 602
    assert FakeEngine.created == created                    # same geometry: the engine is reused, fields re-uploaded


def test_engine_is_rebuilt_when_the_scene_geometry_changes(S):
    s = _quiet(S.Solver, 10, 0.1, 1, 1, 1.0, 2e-9)
    _quiet(s.add_gaussian_pulse, 1e9, (0.5, 0.5))
    s.create_material()
    _quiet(s.solve, solve_only=True)
    n0, created = s.size, FakeEngine.created
    s.set_reflect_boundaries(up=True, down=False, left=False, right=False)   # walls are part of the engine key
    s.reset()
    _quiet(s.solve, solve_only=True)
    assert FakeEngine.created == created + 1 and s.size == n0


def test_stale_pulse_index_error_and_partial_progress(S):
    s = _quiet(S.Solver, 10, 0.1, 1, 1, 1.0, 4e-9)
    _quiet(s.add_gaussian_pulse, 2e9, (0.3, 0.3))
    n_short = len(s.pulses[0]._magnitude_vector)
    _quiet(s.add_oscillating_pulse, 3e9, (0.5, 0.5), 8e9)    # re-meshes: the first pulse keeps its short table
    s.create_material()
    with pytest.raises(IndexError):
        _quiet(s.solve, solve_only=True)
    assert s.step == n_short


def test_second_solve_without_reset_replays_the_reference_literally(S):
    """solve() twice: times restart at 0 while self.step keeps counting (solver.py:284-291)."""
    sc = SC.spec("tutorial")
    s = _quiet(SC.drive, S, sc)
    s.end_time = 120.5 * s.dt
    f, setup, times, _ = SC.restate(sc)
    times = times[:121]
    _quiet(s.solve, solve_only=True)
    R.run(f, setup, times)
    assert np.array_equal(s.h, f.h)
    _quiet(s.solve, solve_only=True)
    for t in times:                                         # the oracle with the same quirk: step continues, time restarts
        R.leapfrog(f, setup, t)
    assert s.step == f.step == 242 and np.array_equal(s.h, f.h)


def test_realtime_path_updates_once_per_frame(S, monkeypatch):
    """solve() without save(): plot_and_show drives one update() per animation frame (solver.py:309-329).  matplotlib
    is replaced by a stand-in whose FuncAnimation simply plays every frame."""
    import sys
    import types
    shown = []

    class Image:
        def set_array(self, a):
            shown.append(np.array(a))

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

    sc = SC.spec("lens")
    f, setup, times, _ = SC.restate(sc)
    s = _quiet(SC.drive, S, sc)
    s.end_time = 60.5 * s.dt
    _quiet(s.solve)                                          # realtime=True and no save file -> plot_and_show
    for t in times[:61]:
        R.leapfrog(f, setup, t)
    assert len(shown) == 61 and s.step == 61 and np.array_equal(shown[-1], f.h)

    # frame_stride = 7: seven calls per frame in one device run, the same fields at the frames shown
    del shown[:]
    f7, setup7, _, _ = SC.restate(sc)
    s7 = _quiet(SC.drive, S, sc)
    s7.end_time = 60.5 * s7.dt
    s7.frame_stride = 7
    _quiet(s7.solve)
    want = []
    for n, t in enumerate(times[:61]):
        R.leapfrog(f7, setup7, t)
        if (n + 1) % 7 == 0 or n == 60:
            want.append(f7.h.copy())
    assert s7.step == 61 and len(shown) == len(want) == 9
    assert all(np.array_equal(a, b) for a, b in zip(shown, want))

    # view_decimation = 3: the same run, every third row / column in the picture; the simulation itself is untouched.
    # With energies recorded: the records of the whole view are collected when it ends
    del shown[:]
    f3, setup3, _, _ = SC.restate(sc)
    s3 = _quiet(SC.drive, S, sc)
    s3.end_time = 60.5 * s3.dt
    s3.frame_stride, s3.view_decimation = 7, 3
    s3.record_energy((1.0, 0.5))
    _quiet(s3.solve)
    assert s3.step == 61 and len(shown) == 9 and all(np.array_equal(a, b[::3, ::3]) for a, b in zip(shown, want))
    assert np.array_equal(s3.h, want[-1]) and len(s3.recorded_energies[0]["energies"]) == 61

    # the window is closed after two frames: the batch that was running ahead is accounted for, nothing is lost
    del shown[:]
    def two_frames(fig, func, frames, **k):
        for frame in list(frames)[:2]:
            func(frame)
    monkeypatch.setattr(anim, "FuncAnimation", two_frames)
    s2 = _quiet(SC.drive, S, sc)
    s2.end_time = 60.5 * s2.dt
    s2.frame_stride = 7
    _quiet(s2.solve)
    assert len(shown) == 2 and s2.step == 21                 # two shown + the one in flight
    f2, setup2, _, _ = SC.restate(sc)
    for t in times[:21]:
        R.leapfrog(f2, setup2, t)
    assert np.array_equal(s2.h, f2.h)


def test_big_snapshot_stacks_stream_into_the_npy_file(S, monkeypatch, tmp_path):
    """above STREAM_SNAPSHOT_BYTES the stack is an open_memmap of <name>.npy, never a RAM copy +
    np.save; the file, h_list and what FileLoader sees are the same as on the in-RAM route."""
    from fdtd_em_solver_b200 import analyser
    sc = SC.spec("tutorial")
    out = {}
    for mode, limit in (("ram", 1 << 40), ("stream", 1)):
        monkeypatch.setattr(S, "STREAM_SNAPSHOT_BYTES", limit)
        s = _quiet(SC.drive, S, sc)
        s.end_time = 300.5 * s.dt
        s.save(str(tmp_path / mode))
        _quiet(s.solve, realtime=False, step_frequency=5)
        assert isinstance(s._stack, np.memmap) == (mode == "stream") and s._streamed is None
        out[mode] = (np.load(str(tmp_path / mode) + ".npy"), np.array(s.h_list))
    assert out["ram"][0].shape == (60, 255, 255)
    assert np.array_equal(out["ram"][0], out["stream"][0]) and np.array_equal(out["ram"][1], out["stream"][1])
    loaded = _quiet(analyser.FileLoader, str(tmp_path / "stream"))
    assert np.array_equal(loaded.get_matrix(), out["ram"][0]) and loaded.size == 255


def test_streamed_file_is_removed_when_the_loop_raises(S, monkeypatch, tmp_path):
    """The reference reaches np.save only after a complete loop (solver.py:297-303): a stale pulse leaves no file."""
    import os
    monkeypatch.setattr(S, "STREAM_SNAPSHOT_BYTES", 1)
    s = _quiet(S.Solver, 10, 0.1, 1, 1, 1.0, 4e-9)
    _quiet(s.add_gaussian_pulse, 2e9, (0.3, 0.3))
    _quiet(s.add_oscillating_pulse, 3e9, (0.5, 0.5), 8e9)
    s.create_material()
    s.save(str(tmp_path / "broken"))
    with pytest.raises(IndexError):
        _quiet(s.solve, realtime=False)
    assert not os.path.exists(str(tmp_path / "broken") + ".npy") and s._streamed is None


def test_update_reflect_by_hand_matches_the_reference(S):
    """solver.py:164-167 is public in the reference; called between steps it zeroes h inside the reflectors."""
    sc = SC.spec("waveguide")
    f, setup, times, info = SC.restate(sc)
    s = _quiet(SC.drive, S, sc)
    for t in times[:30]:
        R.leapfrog(f, setup, t)
        s.update(t)
    s.h = s.h + 1.0                                         # something visible inside the reflectors
    f.h = f.h + 1.0
    s.update_reflect()
    for top, bottom, left, right in s.reflectors:
        f.h[top:bottom, left:right] = 0
        assert not s.h[top:bottom, left:right].any()
    assert np.array_equal(s.h, f.h)
    t = times[30]
    R.leapfrog(f, setup, t)                                 # the edited h is what the next step starts from
    assert np.array_equal(s.update(t), f.h)


def test_public_surface_matches_the_reference_signatures():
    """every public method of the reference's classes exists here with the same leading parameters
    (extra keywords with defaults are allowed).  Needs /root/reference (build container only)."""
    import inspect
    from oracle import ref_import
    if not ref_import.available():
        pytest.skip("reference sources not present on this box")
    ref_solver, ref_analyser = ref_import.load()[:2]
    from fdtd_em_solver_b200 import solver, analyser
    for ref_mod, mod in ((ref_solver, solver), (ref_analyser, analyser)):
        for cname, cls in inspect.getmembers(ref_mod, inspect.isclass):
            if cls.__module__ != ref_mod.__name__:
                continue
            mine = getattr(mod, cname)
            for name, fn in inspect.getmembers(cls, inspect.isfunction):
                theirs = list(inspect.signature(fn).parameters.values())
                ours = list(inspect.signature(getattr(mine, name)).parameters.values())
                assert [(p.name, p.default) for p in ours[:len(theirs)]] == [(p.name, p.default) for p in theirs], \
                    (cname, name)
                assert all(p.default is not inspect.Parameter.empty for p in ours[len(theirs):]), (cname, name)
    for const in ("C", "MU", "EPSILON"):
        assert getattr(solver, const) == getattr(ref_solver, const)


@pytest.mark.parametrize("name,n_calls", [("tutorial", 700), ("waveguide", 600)])
def test_record_field_equals_the_data_collector_of_a_saved_run(S, name, n_calls, tmp_path):
    """the on-device probe series == DataCollector over the saved snapshot stack, bit for bit --
    also AT the point source, where a saved snapshot shows the NEXT call's source write (App. A-Q8)."""
    from fdtd_em_solver_b200 import analyser
    sc = SC.spec(name)
    times = SC.restate(sc)[2][:n_calls]

    def driven():
        s = _quiet(SC.drive, S, sc)
        s.end_time = float(times[-1]) + 0.5 * s.dt
        return s

    saved = driven()
    saved.save(str(tmp_path / name))
    _quiet(saved.solve, realtime=False, step_frequency=sc["step_frequency"])
    loader = _quiet(analyser.FileLoader, str(tmp_path / name))

    probed = driven()
    src = probed.pulses[0]
    at_source = ((src.get_row() + 0.5) * probed.ds, (src.get_col() + 0.5) * probed.ds)
    assert (int(at_source[0] / probed.ds), int(at_source[1] / probed.ds)) == (src.get_row(), src.get_col())
    places = [at_source, (0.4 * probed.length_x, 0.6 * probed.length_y), (0.0, 0.0)]
    for loc in places:
        probed.record_field(loc)
    _quiet(probed.solve, solve_only=True, step_frequency=sc["step_frequency"])
    assert probed.h_list == [] and np.array_equal(probed.h, saved.h)
    some_aliased = False
    for k, loc in enumerate(places):
        want = loader.create_data_collector(loc)
        got = probed.field_collector(k)
        assert got.data.shape == want.data.shape and np.array_equal(got.data, want.data), (name, loc)
        assert np.array_equal(got.time_vector, want.time_vector) and (got.i_pos, got.j_pos) == (want.i_pos, want.j_pos)
        got.fft(); want.fft()
        assert np.array_equal(got.fft_amplitude, want.fft_amplitude)
        assert np.array_equal(got.fft_frequencies, want.fft_frequencies)
    # the source cell really exercises the aliasing rule: its saved samples are pulse magnitudes, not computed h
    d = probed.recorded_fields[0]["samples"]
    some_aliased = any(d[m] == src.magnitude((m + 1) * sc["step_frequency"]) and d[m] != 0 for m in range(len(d) - 1))
    assert some_aliased
    if probed.recorded_energies:                           # energy probes and field probes share the device probe list
        assert probed.recorded_energies[0]["energies"] == saved.recorded_energies[0]["energies"]
