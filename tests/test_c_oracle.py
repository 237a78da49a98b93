"""CPU: the C restatement (oracle/restated.c through oracle/c_oracle.py) is pinned the same two ways as the numpy one:
bitwise against the golden vectors the REAL reference produced (tests/golden/), and bitwise against
oracle/restated.py on seeded random cases with every quirk switched on."""
import copy

import numpy as np
import pytest

from oracle import c_oracle as CO
from oracle import restated as R
from oracle import scenarios as SC
from golden_util import load, random_files, setup_of, sha

pytestmark = pytest.mark.skipif(not CO.available(), reason="gcc not found")


def test_builds_without_fma():
    """The shared object exists, exports oracle_step, and its object code holds no fused multiply-add."""
    import subprocess
    path = CO.build()
    assert hasattr(CO.lib(), "oracle_step")
    asm = subprocess.run(["objdump", "-d", path], capture_output=True, text=True)
    if asm.returncode == 0:
        assert "vfmadd" not in asm.stdout and "vfmsub" not in asm.stdout and "vfnmadd" not in asm.stdout


@pytest.mark.parametrize("path", random_files())
def test_golden_random_cases_every_step(path):
    g = load(path)
    s = setup_of(g)
    f = R.Fields(g["h0"].copy(), g["ex0"].copy(), g["ey0"].copy())
    for n, t in enumerate(g["times"]):
        CO.leapfrog(f, s, t)
        assert np.array_equal(f.h, g["h_steps"][n]), n
    assert np.array_equal(f.h, g["h"]) and np.array_equal(f.ex, g["ex"]) and np.array_equal(f.ey, g["ey"])


@pytest.mark.parametrize("name", ["lens", "mixed", "solver_test", "waveguide_mixed_walls"])
def test_golden_scenarios_full_runs(name):
    g = load(name)
    s = setup_of(g, g["eps"], g["mu"])
    n = int(g["size"])
    f = R.Fields.zeros(n)
    f.probes = [(int(i), int(j), []) for i, j in g["probe_cells"]]
    stack = np.array(CO.run_saving(f, s, g["times"], int(g["step_frequency"])))
    assert len(stack) == int(g["n_snapshots"]) and sha(stack) == str(g["sha_stack"])
    assert np.array_equal(f.h, g["final_h"]) and sha(f.ex) == str(g["sha_ex"]) and sha(f.ey) == str(g["sha_ey"])
    for k, (_, _, series) in enumerate(f.probes):
        assert np.array_equal(np.array(series), g["energies"][k])


WALLS = [(False, False, False, False), (True, True, True, True), (True, False, False, True), (False, True, True, False)]


@pytest.mark.parametrize("seed", range(8))
def test_matches_numpy_restatement(seed):
    rows, cols = [(9, 9), (17, 33), (40, 12), (64, 64), (5, 70), (31, 31), (128, 96), (2, 2)][seed]
    if min(rows, cols) < 3:
        f, s, times = SC.random_case(rows + 2, cols + 2, seed, WALLS[seed % 4], with_sources=False, with_rects=False)
    else:
        f, s, times = SC.random_case(rows, cols, seed, WALLS[seed % 4], n_steps=14)
    fa, fb = copy.deepcopy(f), copy.deepcopy(f)
    fa.raw_probes = [(0, 0, []), (fa.h.shape[0] - 1, fa.h.shape[1] - 1, [])]
    fb.raw_probes = copy.deepcopy(fa.raw_probes)
    for n, t in enumerate(times):
        old_a, old_b = fa.h, fb.h
        R.leapfrog(fa, s, t)
        CO.leapfrog(fb, s, t)
        assert np.array_equal(fa.h, fb.h) and np.array_equal(fa.ex, fb.ex) and np.array_equal(fa.ey, fb.ey), n
        assert np.array_equal(old_a, old_b), n                      # hard point writes into the OLD array (:180)
    assert fa.raw_probes == fb.raw_probes and fa.step == fb.step


def test_snapshot_aliasing_and_stale_pulse():
    """Q8: a saved snapshot is the array the next call's point source writes into.  Q10: a pulse indexed past its
    table raises IndexError, in both restatements at the same step."""
    f, s, times = SC.random_case(24, 24, 3, n_steps=20)
    fa, fb = copy.deepcopy(f), copy.deepcopy(f)
    a = np.array(R.run_saving(fa, s, times, 3))
    b = np.array(CO.run_saving(fb, s, times, 3))
    assert a.shape == b.shape and np.array_equal(a, b)
    short = copy.deepcopy(s)
    short.pulses[0].table = short.pulses[0].table[:5]
    short.pulses[0].start, short.pulses[0].end = -1.0, 1.0
    for step in (R.leapfrog, CO.leapfrog):
        g = copy.deepcopy(f)
        with pytest.raises(IndexError):
            for t in times:
                step(g, short, t)
        assert g.step == 5


def test_thread_count_does_not_change_bits():
    f, s, times = SC.random_case(96, 80, 11, n_steps=6)
    out = []
    for threads in (1, 3):
        CO.set_threads(threads)
        g = copy.deepcopy(f)
        CO.run(g, s, times)
        out.append((g.h, g.ex, g.ey))
    CO.set_threads(1)
    for x, y in zip(*out):
        assert np.array_equal(x, y)
