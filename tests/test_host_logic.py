"""CPU: host-side logic of the product (no GPU, no compute calls through the C ABI):
the drop-in Solver/MaterialArray/Pulse setup against the golden fixtures from the real reference,
the source-table builder, the restated setup functions, and the C-ABI library's exports."""
import contextlib
import ctypes
import io
import os
import re

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC
from golden_util import SCENARIOS, load, pulses_of

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _product(name):
    from fdtd_em_solver_b200 import solver as S
    with contextlib.redirect_stdout(io.StringIO()):
        return SC.drive(S, SC.spec(name))


@pytest.mark.parametrize("name", SCENARIOS)
def test_product_setup_matches_reference_golden(name):
    g = load(name)
    s = _product(name)
    assert s.size == int(g["size"]) and s.dt == float(g["dt"]) and s.ds == float(g["ds"])
    assert s.omega_max == float(g["omega_max"]) and s.length_x == float(g["length"])
    assert np.array_equal(s.material.eps_arr, g["eps"]) and np.array_equal(s.material.mu_arr, g["mu"])
    assert [bool(b) for b in s.boundaries] == [bool(b) for b in g["walls"]]
    assert np.array_equal(np.array(s.reflectors, dtype=np.int64).reshape(-1, 4), g["reflectors"])
    assert [(d["i"], d["j"]) for d in s.recorded_energies] == [tuple(map(int, c)) for c in g["probe_cells"]]
    assert len(s.pulses) == int(g["n_pulses"])
    for p, q in zip(s.pulses, pulses_of(g)):
        assert (p.get_direction(), p.get_row(), p.get_col()) == (q.direction, q.row, q.col)
        assert (p.get_start_time(), p.get_end_time()) == (q.start, q.end)
        assert np.array_equal(p._magnitude_vector, q.table)          # same expressions, same association order
    s._coefficients()
    assert np.array_equal(s.eps_coefficient_arr, g["eps_coef"]) and np.array_equal(s.mu_coefficient_arr, g["mu_coef"])
    assert s.h.shape == (s.size, s.size) and s.ex.shape == (s.size + 1, s.size) and s.ey.shape == (s.size, s.size + 1)
    assert not s.h.any()


@pytest.mark.parametrize("name", SCENARIOS)
def test_restated_setup_matches_reference_golden(name):
    g = load(name)
    f, setup, times, info = SC.restate(SC.spec(name))
    assert info["size"] == int(g["size"]) and info["dt"] == float(g["dt"]) and info["ds"] == float(g["ds"])
    assert np.array_equal(setup.eps, g["eps"]) and np.array_equal(setup.mu, g["mu"])
    full = R.time_axis(info["end_time"], info["dt"])
    assert np.array_equal(full[:len(g["times"])], g["times"])
    for p, q in zip(setup.pulses, pulses_of(g)):
        assert np.array_equal(p.table, q.table) and (p.row, p.col, p.start, p.end) == (q.row, q.col, q.start, q.end)


def test_source_tables_follow_list_order_and_strict_gating():
    from fdtd_em_solver_b200.engine import build_source_tables, IMPEDANCE
    times = np.arange(10) * 1.0
    A = R.PulseSpec("t", None, 1, 2, 0, 5.0, np.arange(10) + 100.0)          # active for 0 < t < 5
    B = R.PulseSpec("t", "left", 3, 4, 2.0, 8.0, np.arange(10) + 200.0)      # active for 2 < t < 8
    Cc = R.PulseSpec("t", "up", 0, 0, 6.0, 7.5, np.arange(6) + 300.0)        # short table, active only at t=7
    srcs, mag, magz, active, last, first_bad = build_source_tables([A, B, Cc], times)
    assert active[0].tolist() == [0, 1, 1, 1, 1, 0, 0, 0, 0, 0]               # strict compare: t=0 and t=5 excluded
    assert active[1].tolist() == [0, 0, 0, 1, 1, 1, 1, 1, 0, 0]
    assert active[2].tolist() == [0, 0, 0, 0, 0, 0, 0, 1, 0, 0]
    assert first_bad == 7                                                     # table of length 6 indexed at 7
    assert last[1] == 101 and last[3] == 203 and last[5] == 205               # last ACTIVE pulse in list order
    assert np.array_equal(magz, mag * IMPEDANCE)
    assert [(s.kind, s.row, s.col) for s in srcs] == [(0, 1, 2), (2, 3, 4), (3, 0, 0)]


def test_material_convex_matches_the_cellwise_reference_semantics():
    """Vectorised set_material_convex vs the restated cell-by-cell loops (negative indices wrap, overflow skipped)."""
    from fdtd_em_solver_b200.solver import MaterialArray
    for size, length, center, radius, thickness in [(60, 3.0, (1.5, 1.0), 2.2, 1.2), (40, 2.0, (0.2, 0.3), 1.0, 0.5),
                                                    (33, 1.0, (0.9, 0.9), 0.6, 0.3)]:
        m = MaterialArray(size, length, length)
        m.set_material_convex(center, radius=radius, thickness=thickness, epsilon_rel=5, mu_rel=2)
        eps, mu = R.vacuum(size)
        R.paint_convex(eps, mu, center, radius, thickness, 5, 2, length, size)
        assert np.array_equal(m.eps_arr, eps) and np.array_equal(m.mu_arr, mu)


def test_save_file_formats(tmp_path):
    """save_to_json keys/values and FileLoader/DataCollector round trip (reference solver.py:67-91, analyser.py)."""
    import json
    from fdtd_em_solver_b200 import analyser as A
    s = _product("waveguide")
    s.step_frequency = 5
    s.save(str(tmp_path / "w"))
    s.save_to_json()
    d = json.load(open(str(tmp_path / "w.txt")))
    assert list(d.keys()) == ['dt', 'ds', 'length_x', 'length_y', 'stability', 'omega_max', 'end_time',
                              'step_frequency', 'material', 'pulses']
    assert d["pulses"] == [{'sigma_w': 5 * 10 ** 8, 'omega_0': 2 * 10 ** 9, 'omega_max': 3.5e9,
                            'location': [0.7, 0], 'type': 'oscillate'}]
    stack = np.random.default_rng(0).standard_normal((7, s.size, s.size))
    np.save(str(tmp_path / "w"), stack)
    with contextlib.redirect_stdout(io.StringIO()):
        fl = A.FileLoader(str(tmp_path / "w"))
    assert fl.size == s.size and fl.dt == s.dt and fl.step_frequency == 5 and fl.convert(1.0) == int(1.0 / s.ds)
    dc = fl.create_data_collector((1.0, 2.0))
    assert np.array_equal(dc.data, stack[:, int(1.0 / s.ds), int(2.0 / s.ds)])
    dc.fft()
    assert len(dc.fft_amplitude) == len(dc.fft_frequencies) > 0 and (dc.fft_frequencies < s.omega_max).all()


def test_reference_error_messages():
    from fdtd_em_solver_b200 import solver as S
    with contextlib.redirect_stdout(io.StringIO()):
        s = S.Solver(10, 0.1, 1, 1, 1.0, 1e-9)
        with pytest.raises(Exception, match="No pulse added"):
            s.solve()
        s.add_gaussian_pulse(1e9, (0.5, 0.5))
        s.create_material()
        with pytest.raises(Exception, match="Must create a save file name"):
            s.solve(realtime=False)
        with pytest.raises(Exception, match="Pulse type not recognised"):
            S.Pulse(1e-12, 0, "nope", 1e-9, sigma_w=1e9, location=[1, 1])
        with pytest.raises(Exception, match="Pulse direction not recognised"):
            S.Pulse(1e-12, 0, "gd", 1e-9, sigma_w=1e9, location=[1, 1], direction="sideways")


def test_c_abi_library_exports_every_declared_symbol():
    """The shared library loads without a GPU and exports exactly what include/fdtd2d.h declares."""
    from fdtd_em_solver_b200 import _lib as L
    header = open(os.path.join(ROOT, "include", "fdtd2d.h")).read()
    declared = set(re.findall(r"\b(fdtd_[a-z0-9_]+)\s*\(", header))
    assert declared == set(L.SIGNATURES), declared ^ set(L.SIGNATURES)
    lib = L.load()
    for name in declared:
        assert hasattr(lib, name), name
    assert lib.fdtd_default_pitch(65536, 0) == 65568 and lib.fdtd_default_pitch(255, 1) == 256
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = 100, 50, 10, 30
    assert lib.fdtd_local_rows(ctypes.byref(cfg)) == 22


def test_no_gpu_means_loud_failure_not_a_cpu_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    from fdtd_em_solver_b200.engine import Engine
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        Engine(16, 16, courant=0.1)
    s = _product("lens")
    with contextlib.redirect_stdout(io.StringIO()):
        with pytest.raises(RuntimeError, match="no CPU fallback"):
            s.solve(solve_only=True)


def test_recorded_shapes_replay_and_cell_lookup():
    """MaterialArray(on_device=True): the recorded shapes reproduce the eager host arrays (lazy materialisation and
    the single-cell lookup used by the energy probes)."""
    from fdtd_em_solver_b200.solver import MaterialArray
    def paint(m):
        m.set_material_rect((0.2, 0.3), (0.9, 0.7), 3, 2)
        m.set_material_convex((0.1, 0.2), radius=0.9, thickness=0.5, epsilon_rel=5)          # wraps negative indices
        with contextlib.redirect_stdout(io.StringIO()):
            m.set_fixed_length_waveguide((0.5, 0), 1.2, 0.1, 0.1, 2.7)
            m.set_quarter_circle((0.6, 0.4), 0.3, 2.5, mu_rel=2, thickness=0.1)
        m.set_material_rect((0.8, 0.8), (1.0, 1.0), 4)                                       # far edge: convert -> size
        m.set_material_rect((5, 40), (200, 300), 1.5, convert=False)                         # clipped by slicing
    eager, lazy = MaterialArray(57, 1.0, 1.0), MaterialArray(57, 1.0, 1.0, on_device=True)
    paint(eager); paint(lazy)
    assert not lazy.on_host and lazy.ops == eager.ops and len(lazy.ops) == 8
    cells = [(i, j) for i in range(57) for j in range(57)]
    got = np.array([lazy.cell(i, j) for i, j in cells])
    assert not lazy.on_host
    assert np.array_equal(got[:, 0].reshape(57, 57), eager.eps_arr) and np.array_equal(got[:, 1].reshape(57, 57), eager.mu_arr)
    assert np.array_equal(lazy.eps_arr, eager.eps_arr) and np.array_equal(lazy.mu_arr, eager.mu_arr) and lazy.on_host


def test_c_abi_argument_validation_needs_no_gpu():
    """Geometry checks of fdtd_create happen before any device is touched: error codes + messages (include/fdtd2d.h)."""
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()

    def create(**kw):
        cfg = L.Config()
        cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = 64, 64, 0, 64
        cfg.kernel = L.KERNEL_STREAM
        for k, v in kw.items():
            setattr(cfg, k, v)
        h = ctypes.c_void_p()
        rc = lib.fdtd_create(ctypes.byref(cfg), ctypes.byref(h))
        msg = lib.fdtd_last_error(None).decode()
        if rc == 0:
            lib.fdtd_destroy(h)
        return rc, msg

    assert create(rows=2)[0] == -1 and "at least 3x3" in create(rows=2)[1]
    assert create(cols=1)[0] == -1
    assert create(row_begin=10, row_end=5)[0] == -1 and "slab" in create(row_begin=10, row_end=5)[1]
    assert create(row_begin=0, row_end=1)[0] == -1 and "at least 2 rows" in create(row_begin=0, row_end=1)[1]
    assert create(row_end=65)[0] == -1
    assert create(dtype=7)[0] == -1 and "dtype" in create(dtype=7)[1]
    assert create(kernel=9)[0] == -1
    assert lib.fdtd_create(None, None) == -1
    # NULL handles are rejected, never dereferenced
    assert lib.fdtd_step(None, 0) == -1 and lib.fdtd_sync(None) == -1 and lib.fdtd_destroy(None) == 0
    assert lib.fdtd_set_temporal_blocking(None, 2) == -1 and lib.fdtd_current_generation(None) == -1
    assert lib.fdtd_set_resident(None, 1, 0) == -1 and lib.fdtd_resident_active(None) == -1
    assert lib.fdtd_set_frame_mode(None, 3) == -1 and lib.fdtd_get_tuning(None, None) == -1
    assert lib.fdtd_uniform_mu(None) == -1 and lib.fdtd_rescan_coeffs(None) == -1


def test_snapshot_stack_allocator_is_a_plain_float64_array():
    """engine.empty_hugepages: np.empty + a MADV_HUGEPAGE hint (only a hint: never an error, never a different array)."""
    from fdtd_em_solver_b200.engine import empty_hugepages
    for shape in ((3, 5, 7), (40, 255, 255), (0, 4, 4)):
        a = empty_hugepages(shape)
        assert a.shape == shape and a.dtype == np.float64 and a.flags.c_contiguous and a.flags.writeable
        if a.size:
            a[...] = 1.5
            assert float(a.sum()) == 1.5 * a.size


def test_integration_md_stub_compiles_and_matches_the_abi():
    """INTEGRATION.md's reference-side binding must not drift from include/fdtd2d.h / _lib.py: its code block compiles,
    its ctypes structures have the library's field lists, and every fdtd_* symbol it calls is exported."""
    import ast
    import re
    from fdtd_em_solver_b200 import _lib as L
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    blocks = re.findall(r"```python\n(.*?)```", text, flags=re.S)
    stub = max(blocks, key=len)
    tree = ast.parse(stub)                                     # compiles
    fields = {}
    for node in ast.walk(tree):
        if isinstance(node, ast.ClassDef):
            for stmt in node.body:
                if isinstance(stmt, ast.Assign) and stmt.targets[0].id == "_fields_":
                    fields[node.name] = [elt.elts[0].value for elt in stmt.value.elts]
    assert fields["Config"] == [f[0] for f in L.Config._fields_]
    assert fields["Source"] == [f[0] for f in L.Source._fields_]
    called = set(re.findall(r"lib\.(fdtd_\w+)", stub))
    assert called and called <= set(L.SIGNATURES), called - set(L.SIGNATURES)
    direction_kinds = ast.literal_eval(re.search(r"KIND = (\{.*?\})", stub).group(1))
    assert direction_kinds == L.DIRECTION_KIND
