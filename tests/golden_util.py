"""Loading of tests/golden/*.npz (written by oracle/make_golden.py from the real reference)."""
import glob
import hashlib
import os

import numpy as np

from oracle import restated as R

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
DIRS = {0: None, 1: "right", 2: "left", 3: "up", 4: "down"}
SCENARIOS = ["tutorial", "doubleslit", "lens", "waveguide", "waveguide_mixed_walls", "mixed", "solver_test"]


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def random_files():
    return sorted(glob.glob(os.path.join(GOLDEN, "random_*.npz")))


def load(name):
    path = name if os.path.isabs(name) else os.path.join(GOLDEN, "scenario_%s.npz" % name)
    return np.load(path, allow_pickle=False)


def pulses_of(g):
    out = []
    for k in range(int(g["n_pulses"])):
        r, c = (int(v) for v in g["pulse_cell"][k])
        t0, t1 = (float(v) for v in g["pulse_window"][k])
        out.append(R.PulseSpec("table", DIRS[int(g["pulse_dir"][k])], r, c, t0, t1, g["pulse_table_%d" % k]))
    return out


def setup_of(g, eps=None, mu=None):
    return R.Setup(eps_coef=g["eps_coef"], mu_coef=g["mu_coef"], courant=float(g["courant"]),
                   reflect_walls=tuple(bool(v) for v in g["walls"]), pulses=pulses_of(g),
                   reflectors=[list(map(int, r)) for r in g["reflectors"]], eps=eps, mu=mu)
