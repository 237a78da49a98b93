"""TEST INFRASTRUCTURE ONLY -- the BASELINE.json configs as data.

Each scenario is the sequence of public-API calls the reference script makes
(repaired as SURVEY.md App. C describes), so the same spec can drive

  * the real reference (`drive(ref_solver_module, spec)`, build container only),
  * the product (`drive(fdtd_em_solver_b200.solver, spec)`), and
  * the numpy restatement (`restate(spec)`), which re-derives everything
    with oracle/restated.py and no Solver object at all.

Sources: tutorial.py:5-36, doubleslit.py:5-38, lens.py:4-21, waveguide.py:6-40,
solver.py:645-665 (`test()`); the "mixed" case is synthetic.
"""
import numpy as np

from . import restated as R

C = R.C


def _tutorial():
    return dict(
        ctor=dict(points_per_wavelength=10, stability=0.2, eps_r_max=4, mu_r_max=1,
                  simulation_size=3, simulation_time=2.5 * 10 ** (-8)),
        calls=[
            ("add_oscillating_pulse", (1 * 10 ** 9, (0.8, 0.4), 5 * 10 ** 9), {}),
            ("create_material", (), {}),
            ("set_reflect_boundaries", (), dict(up=False, down=False, left=False, right=False)),
        ],
        step_frequency=5, probe_cells=[(127, 127)])


def _doubleslit():
    size = 4
    return dict(
        ctor=dict(points_per_wavelength=10, stability=1 / 10, eps_r_max=1, mu_r_max=1,
                  simulation_size=size, simulation_time=10 * size / C),
        calls=[
            ("add_oscillating_pulse", (3.0 * 10 ** 9, (2, 0.1), 6 * 10 ** 9), dict(direction="right")),
            ("create_material", (), {}),
            ("set_reflect_square", ((0, 1), (1.62, 1.05)), {}),
            ("set_reflect_square", ((1.76, 1), (2.24, 1.05)), {}),
            ("set_reflect_square", ((2.38, 1), (size, 1.05)), {}),
            ("set_reflect_boundaries", (), dict(up=False, down=False, left=False, right=False)),
        ],
        step_frequency=5, probe_cells=[(159, 300)])


def _lens():
    return dict(
        ctor=dict(points_per_wavelength=10, stability=0.1, eps_r_max=5, mu_r_max=1,
                  simulation_size=3, simulation_time=2 * 10 ** (-8)),
        calls=[
            ("add_sinusoidal_pulse", (), dict(location=(1.5, 0.2), omega_0=5 * 10 ** 9, direction="right")),
            ("create_material", (), {}),
            ("material.set_material_convex", ((1.5, 1),), dict(radius=2.2, thickness=1.2, epsilon_rel=5)),
            ("set_reflect_boundaries", (), dict(up=False, down=False, right=False, left=False)),
        ],
        step_frequency=5, probe_cells=[(89, 120)])


def _waveguide(mixed_walls=False):
    size = 4
    calls = [
        ("set_reflect_boundaries", (False, False, False, False), {}),
        ("add_oscillating_pulse", (5 * 10 ** 8, (0.7, 0), 2 * 10 ** 9), {}),
        ("create_material", (), {}),
        ("material.set_fixed_length_waveguide", ((0.5, 0), 6, 0.1, 0.4, 2.7), {}),
        ("set_reflect_square", ((0, 0.25), (0.5, 0.35)), {}),
        ("set_reflect_square", ((0.9, 0.25), (4, 0.35)), {}),
        ("record_energy", ((0.7, 0.1),), {}),
    ]
    if mixed_walls:
        calls.append(("set_reflect_boundaries", (), dict(up=True, down=False, left=True, right=False)))
    return dict(
        ctor=dict(points_per_wavelength=10, stability=1 / 10, eps_r_max=3.5, mu_r_max=1,
                  simulation_size=size, simulation_time=4 * size / C),
        calls=calls, step_frequency=5, probe_cells=[(24, 3)])


def _solver_test():
    return dict(
        ctor=dict(points_per_wavelength=10, stability=0.1, eps_r_max=4, mu_r_max=3,
                  simulation_size=3, simulation_time=2.5 * 10 ** (-8)),
        calls=[
            ("add_oscillating_pulse", (1 * 10 ** 9, (0.8, 0.4), 5 * 10 ** 9), {}),
            ("create_material", (), {}),
            ("material.set_material_rect", ((2, 2), (2.8, 2.8), 3, 3), {}),
        ],
        step_frequency=50, probe_cells=[(300, 300)])


def _mixed():
    """Every pulse direction, mu_r != 1, reflectors touching walls, walls T/F/T/F."""
    return dict(
        ctor=dict(points_per_wavelength=10, stability=0.15, eps_r_max=4, mu_r_max=3,
                  simulation_size=1.5, simulation_time=1.2 * 10 ** (-8)),
        calls=[
            ("add_oscillating_pulse", (1 * 10 ** 9, (0.4, 0.5), 5 * 10 ** 9), {}),
            ("add_gaussian_pulse", (1.5 * 10 ** 9, (0.9, 0.3)), dict(direction="left")),
            ("add_oscillating_pulse", (1 * 10 ** 9, (0.6, 0.7), 4 * 10 ** 9), dict(direction="up", start_time=1e-9)),
            ("add_sinusoidal_pulse", (3 * 10 ** 9, (1.1, 1.0)), dict(direction="down")),
            ("add_oscillating_pulse", (1 * 10 ** 9, (0.2, 0.2), 5 * 10 ** 9), dict(direction="right")),
            ("add_gaussian_pulse", (1 * 10 ** 9, (1.2, 1.2)), dict(start_time=5e-10)),
            ("create_material", (), {}),
            ("material.set_material_rect", ((0.9, 0.9), (1.3, 1.3), 3, 3), {}),
            ("material.set_quarter_circle", ((0.5, 0.4), 0.3, 2.5), dict(mu_rel=2, thickness=0.1)),
            ("set_reflect_square", ((0, 1.0), (0.3, 1.1)), {}),
            ("set_reflect_square", ((1.2, 0.1), (1.5, 0.4)), {}),
            ("set_reflect_boundaries", (), dict(up=True, down=False, left=True, right=False)),
            ("record_energy", ((0.75, 0.75),), {}),
        ],
        step_frequency=5, probe_cells=[(40, 40)])


SCENARIOS = {
    "tutorial": _tutorial,
    "doubleslit": _doubleslit,
    "lens": _lens,
    "waveguide": _waveguide,
    "waveguide_mixed_walls": lambda: _waveguide(True),
    "solver_test": _solver_test,
    "mixed": _mixed,
}


def spec(name):
    return SCENARIOS[name]()


# --------------------------------------------------------------------------
def drive(solver_module, sc, **ctor_extra):
    """Replay the scenario against anything with the reference's Python API."""
    s = solver_module.Solver(**sc["ctor"], **ctor_extra)
    for name, args, kwargs in sc["calls"]:
        target = s
        if name.startswith("material."):
            target, name = s.material, name.split(".", 1)[1]
        getattr(target, name)(*args, **kwargs)
    return s


def restate(sc):
    """Re-derive the scenario with oracle/restated.py only.

    Returns (Fields, Setup, times, info) where info carries the derived grid
    constants.  Mirrors the ORDER effects of the API: a pulse that raises
    omega_max re-meshes (solver.py:127-129) while earlier pulses keep their
    old dt / table / index location."""
    k = sc["ctor"]
    length, end = k["simulation_size"], k["simulation_time"]
    omega_max = 0
    g = None
    pulses, rects, probes = [], [], []
    walls = (False, False, False, False)
    eps = mu = None

    def need(omega):
        nonlocal omega_max, g
        if omega > omega_max:
            omega_max = omega
            g = R.grid_constants(k["points_per_wavelength"], k["stability"], k["eps_r_max"],
                                 k["mu_r_max"], length, end, omega_max)

    def idx(v):
        return R.to_index(v, length, g["size"])

    for name, args, kw in sc["calls"]:
        if name == "add_oscillating_pulse":
            sigma_w, loc, omega_0 = args
            need(3 * sigma_w + omega_0)
            pulses.append(R.pulse_spec("oscillate", g["dt"], end, [idx(v) for v in loc], sigma_w=sigma_w,
                                       omega_0=omega_0, start_time=kw.get("start_time", 0),
                                       direction=kw.get("direction")))
        elif name == "add_gaussian_pulse":
            sigma_w, loc = args
            need(3 * sigma_w)
            pulses.append(R.pulse_spec("gd", g["dt"], end, [idx(v) for v in loc], sigma_w=sigma_w,
                                       start_time=kw.get("start_time", 0), direction=kw.get("direction")))
        elif name == "add_sinusoidal_pulse":
            omega_0 = kw["omega_0"] if "omega_0" in kw else args[0]
            loc = kw["location"] if "location" in kw else args[1]
            need(omega_0)
            pulses.append(R.pulse_spec("sinusoidal", g["dt"], end, [idx(v) for v in loc], omega_0=omega_0,
                                       start_time=kw.get("start_time", 0), direction=kw.get("direction")))
        elif name == "create_material":
            eps, mu = R.vacuum(g["size"])
        elif name == "material.set_material_rect":
            ul, lr, er = args[0], args[1], args[2]
            mr = args[3] if len(args) > 3 else kw.get("mu_rel", 1)
            R.paint_rect(eps, mu, [R.mat_index(v, length, g["size"]) for v in ul],
                         [R.mat_index(v, length, g["size"]) for v in lr], er, mr)
        elif name == "material.set_material_convex":
            R.paint_convex(eps, mu, args[0], kw["radius"], kw["thickness"], kw["epsilon_rel"],
                           kw.get("mu_rel", 1), length, g["size"])
        elif name == "material.set_quarter_circle":
            center, radius, er = args
            m = lambda v: R.mat_index(v, length, g["size"])
            R.paint_quarter_circle(eps, mu, (m(center[0]), m(center[1])), m(radius), er,
                                   kw.get("mu_rel", 1), m(kw.get("thickness", 0)))
        elif name == "material.set_fixed_length_waveguide":
            R.paint_fixed_length_waveguide(eps, mu, args[0], args[1], args[2], args[3], args[4],
                                           kw.get("mu_rel", 1), length, g["size"])
        elif name == "set_reflect_square":
            ul, lr = args
            rects.append([idx(ul[0]), idx(lr[0]), idx(ul[1]), idx(lr[1])])
        elif name == "set_reflect_boundaries":
            names = ("up", "down", "left", "right")
            vals = dict(up=True, down=True, left=True, right=True)
            vals.update(dict(zip(names, args)))
            vals.update(kw)
            walls = tuple(vals[n] for n in names)
        elif name == "record_energy":
            probes.append(tuple(idx(v) for v in args[0]))
        else:
            raise KeyError(name)

    eps_coef, mu_coef = R.coefficients(eps, mu, g["dt"], g["ds"])
    setup = R.Setup(eps_coef=eps_coef, mu_coef=mu_coef, courant=(C * g["dt"] / g["ds"]),
                    reflect_walls=walls, pulses=pulses, reflectors=rects, eps=eps, mu=mu)
    fields = R.Fields.zeros(g["size"])
    fields.probes = [(i, j, []) for i, j in probes]
    times = R.time_axis(end, g["dt"])
    info = dict(g, length=length, end_time=end, omega_max=omega_max)
    return fields, setup, times, info


def random_case(rows, cols, seed, walls=(False, False, False, False), n_steps=12,
                with_sources=True, with_rects=True, hetero=True):
    """Index-level random case for step-by-step parity: random non-zero
    initial fields, heterogeneous eps AND mu, all pulse kinds, reflectors
    touching walls / clipped / overlapping."""
    rng = np.random.default_rng(seed)
    dt, ds = 4.0e-12, 0.0125
    f = R.Fields(rng.standard_normal((rows, cols)), rng.standard_normal((rows + 1, cols)) * 300.0,
                 rng.standard_normal((rows, cols + 1)) * 300.0)
    if hetero:
        eps = (1 + 3 * rng.random((rows, cols))) * R.EPSILON
        mu = (1 + 2 * rng.random((rows, cols))) * R.MU
    else:
        eps = np.ones((rows, cols)) * R.EPSILON
        mu = np.ones((rows, cols)) * R.MU
    epc, muc = R.coefficients(eps, mu, dt, ds)
    times = np.arange(n_steps) * dt
    pulses = []
    if with_sources:
        def tab():
            return rng.standard_normal(n_steps) * 2.0

        def window():
            a = int(rng.integers(0, max(1, n_steps // 3)))
            b = int(rng.integers(a + 2, n_steps + 2))
            return (a - 0.5) * dt, (b - 0.5) * dt

        def cell():
            return int(rng.integers(0, rows)), int(rng.integers(0, cols))
        kinds = [None, "right", "left", "up", "down", None, "right"]
        for d in kinds:
            r, c = cell()
            if d == "up":
                r = int(rng.integers(0, rows - 1))
            a, b = window()
            pulses.append(R.PulseSpec("table", d, r, c, a, b, tab()))
        # point sources on a wall row / column / corner, and a right-line at column 0 (wraps, Q4)
        a, b = window(); pulses.append(R.PulseSpec("table", None, 0, int(rng.integers(0, cols)), a, b, tab()))
        a, b = window(); pulses.append(R.PulseSpec("table", None, rows - 1, cols - 1, a, b, tab()))
        a, b = window(); pulses.append(R.PulseSpec("table", None, int(rng.integers(0, rows)), 0, a, b, tab()))
        if seed % 3 == 0:
            a, b = window(); pulses.append(R.PulseSpec("table", "right", 0, 0, a, b, tab()))
        if seed % 4 == 1:
            a, b = window(); pulses.append(R.PulseSpec("table", "right", 0, cols - 1, a, b, tab()))
    rects = []
    if with_rects:
        def rect():
            r0 = int(rng.integers(0, rows)); c0 = int(rng.integers(0, cols))
            return [r0, r0 + int(rng.integers(1, rows)), c0, c0 + int(rng.integers(1, max(2, cols // 3)))]
        rects = [rect(), rect(), [0, max(1, rows // 3), cols // 2, cols // 2 + 2],
                 [rows - 2, rows + 5, 0, 3], [rows // 2, rows * 40, cols - 1, cols * 70]]
    setup = R.Setup(eps_coef=epc, mu_coef=muc, courant=0.1 + 0.05 * (seed % 5), reflect_walls=tuple(walls),
                    pulses=pulses, reflectors=rects, eps=eps, mu=mu)
    return f, setup, times
