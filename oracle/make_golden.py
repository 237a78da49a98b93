"""TEST INFRASTRUCTURE ONLY -- generates tests/golden/*.npz by running the REAL reference
(/root/reference, imported through oracle/ref_import.py) in the build container.

    python -m oracle.make_golden

Each fixture stores the complete inputs of the hot path exactly as the reference derived them
(grid constants, material arrays, pulse tables and windows, reflectors, walls, probes) and the
reference's outputs (final h in full, SHA-256 of ex / ey / the saved snapshot stack, a probe-cell
time series through the snapshots, energy-probe series).  Because the pulse tables are stored, a
consumer needs no libm call: everything downstream is correctly rounded +,-,*,/ and reproduces
bit-for-bit on any machine.  The GPU box has no /root/reference; it reads these files instead.
"""
import contextlib
import hashlib
import io
import os

import numpy as np

from . import ref_import, scenarios as SC

OUT = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")
TRUNCATE = {"doubleslit": 3000, "solver_test": 600}       # calls; full runs otherwise
# ... and the full-length doubleslit run (31 938 calls, 6 387 snapshots = 5.2 GB stack) as a second fixture that keeps only
# hashes, the final h and one series:  python -m oracle.make_golden full
FULL = ("doubleslit",)
DIR_CODE = {None: 0, "right": 1, "left": 2, "up": 3, "down": 4}


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def scenario_fixture(ref, name, truncate=True):
    sc = SC.spec(name)
    with contextlib.redirect_stdout(io.StringIO()):
        s = SC.drive(ref, sc)
        n_calls = TRUNCATE.get(name) if truncate else None
        if n_calls is not None:
            full = np.arange(0, s.end_time, s.dt)
            s.end_time = float(full[n_calls - 1]) + 0.5 * s.dt
            assert len(np.arange(0, s.end_time, s.dt)) == n_calls
        tmp = os.path.join("/tmp", "golden_" + name)
        s.save(tmp)
        s.solve(realtime=False, step_frequency=sc["step_frequency"])
    stack = np.load(tmp + ".npy")
    os.remove(tmp + ".npy"); os.remove(tmp + ".txt")
    i, j = sc["probe_cells"][0]
    d = dict(
        size=s.size, dt=s.dt, ds=s.ds, steps=s.steps, end_time=s.end_time, omega_max=s.omega_max,
        length=s.length_x, courant=(ref.C * s.dt / s.ds), step_frequency=sc["step_frequency"],
        times=s.time_vector, eps=s.material.eps_arr, mu=s.material.mu_arr,
        eps_coef=s.eps_coefficient_arr, mu_coef=s.mu_coefficient_arr,
        walls=np.array(s.boundaries, dtype=np.int8), reflectors=np.array(s.reflectors, dtype=np.int64).reshape(-1, 4),
        probe_cells=np.array([(r["i"], r["j"]) for r in s.recorded_energies], dtype=np.int64).reshape(-1, 2),
        energies=np.array([r["energies"] for r in s.recorded_energies], dtype=np.float64),
        n_pulses=len(s.pulses),
        pulse_dir=np.array([DIR_CODE[p.get_direction()] for p in s.pulses], dtype=np.int8),
        pulse_cell=np.array([(p.get_row(), p.get_col()) for p in s.pulses], dtype=np.int64),
        pulse_window=np.array([(p.get_start_time(), p.get_end_time()) for p in s.pulses], dtype=np.float64),
        final_h=s.h, sha_ex=sha(s.ex), sha_ey=sha(s.ey), sha_h=sha(s.h),
        n_snapshots=len(stack), sha_stack=sha(stack), series_cell=np.array([i, j]), series=stack[:, i, j].copy(),
        snap_first=stack[0], snap_mid=stack[len(stack) // 3], norms=np.array(
            [np.linalg.norm(s.h), np.linalg.norm(s.ex), np.linalg.norm(s.ey)]),
    )
    for k, p in enumerate(s.pulses):
        d["pulse_table_%d" % k] = p._magnitude_vector
    return d


def random_fixture(ref, rows, cols, seed, walls):
    """Index-level case run through the reference's own update()."""
    f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=12)
    d = dict(rows=rows, cols=cols, courant=setup.courant, walls=np.array(walls, dtype=np.int8), times=times,
             h0=f.h.copy(), ex0=f.ex.copy(), ey0=f.ey.copy(), eps_coef=setup.eps_coef, mu_coef=setup.mu_coef,
             reflectors=np.array(setup.reflectors, dtype=np.int64).reshape(-1, 4), n_pulses=len(setup.pulses),
             pulse_dir=np.array([DIR_CODE[p.direction] for p in setup.pulses], dtype=np.int8),
             pulse_cell=np.array([(p.row, p.col) for p in setup.pulses], dtype=np.int64),
             pulse_window=np.array([(p.start, p.end) for p in setup.pulses], dtype=np.float64))
    for k, p in enumerate(setup.pulses):
        d["pulse_table_%d" % k] = p.table
    with contextlib.redirect_stdout(io.StringIO()):
        s = ref.Solver(10, 0.1, 1, 1, 1.0, 1e-9)
    s.size, s.dt, s.ds = rows, 1.0, ref.C / setup.courant
    d["courant"] = (ref.C * s.dt / s.ds)
    s.h, s.ex, s.ey = f.h.copy(), f.ex.copy(), f.ey.copy()
    s.eps_coefficient_arr, s.mu_coefficient_arr = setup.eps_coef, setup.mu_coef
    s.reflectors = [list(r) for r in setup.reflectors]
    s.reflect = bool(s.reflectors)
    s.boundaries = list(walls)

    class P:
        def __init__(self, p): self.p = p
        get_start_time = lambda self: self.p.start
        get_end_time = lambda self: self.p.end
        get_direction = lambda self: self.p.direction
        get_row = lambda self: self.p.row
        get_col = lambda self: self.p.col
        magnitude = lambda self, n: self.p.table[n]
    s.pulses = [P(p) for p in setup.pulses]
    hs = []
    for t in times:
        s.update(t)
        hs.append(s.h.copy())
    d.update(h_steps=np.array(hs), h=s.h, ex=s.ex, ey=s.ey)
    return d


def main():
    import sys
    ref = ref_import.load()[0]
    os.makedirs(OUT, exist_ok=True)
    if "full" in sys.argv[1:]:
        for name in FULL:
            d = scenario_fixture(ref, name, truncate=False)
            for big in ("snap_first", "snap_mid", "eps", "mu"):              # the truncated fixture carries these already
                d.pop(big, None)
            np.savez_compressed(os.path.join(OUT, "scenario_%s_full.npz" % name), **d)
            print("wrote", name, "full length:", len(d["times"]), "calls,", int(d["n_snapshots"]), "snapshots")
        return
    for name in ("tutorial", "doubleslit", "lens", "waveguide", "waveguide_mixed_walls", "mixed", "solver_test"):
        np.savez_compressed(os.path.join(OUT, "scenario_%s.npz" % name), **scenario_fixture(ref, name))
        print("wrote", name)
    cases = [(5, 5, 0, (False, False, False, False)), (8, 13, 1, (True, False, True, False)),
             (17, 9, 2, (False, True, False, True)), (33, 33, 3, (False, False, True, True)),
             (40, 70, 4, (False, False, False, False)), (3, 3, 5, (False, False, False, False))]
    for rows, cols, seed, walls in cases:
        np.savez_compressed(os.path.join(OUT, "random_%dx%d_s%d.npz" % (rows, cols, seed)),
                            **random_fixture(ref, rows, cols, seed, walls))
    print("wrote", len(cases), "random cases")


if __name__ == "__main__":
    main()
