"""CPU: the example scripts (BASELINE configs 1-4 as a user of the reference would write them) run end to end through
the engine double and leave exactly the oracle's snapshot stack for the same scene -- i.e. the scripts really are the
configs the parity suite checks, not look-alikes."""
import contextlib
import copy
import io
import os
import runpy
import sys

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC
from fake_engine import FakeEngine

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")


@pytest.mark.parametrize("script,scenario,scale,extra", [
    ("tutorial.py", "tutorial", 0.05, []), ("doubleslit.py", "doubleslit", 0.004, []), ("lens.py", "lens", 0.05, []),
    ("waveguide.py", "waveguide", 0.03, []), ("waveguide.py", "waveguide_mixed_walls", 0.03, ["--mixed-walls"])])
def test_example_script_is_the_baseline_config(script, scenario, scale, extra, tmp_path, monkeypatch):
    from fdtd_em_solver_b200 import solver
    monkeypatch.setattr(solver, "Engine", FakeEngine)
    name = str(tmp_path / scenario)
    monkeypatch.setattr(sys, "argv", [script, "--name", name, "--time-scale", repr(scale)] + extra)
    monkeypatch.syspath_prepend(EXAMPLES)
    monkeypatch.delitem(sys.modules, "_common", raising=False)
    out = io.StringIO()
    with contextlib.redirect_stdout(out):
        ns = runpy.run_path(os.path.join(EXAMPLES, script), run_name="__main__")
    sc = copy.deepcopy(SC.spec(scenario))
    sc["ctor"]["simulation_time"] = sc["ctor"]["simulation_time"] * scale
    f, setup, times, info = SC.restate(sc)
    want = np.array(R.run_saving(f, setup, times, sc["step_frequency"]))
    got = np.load(name + ".npy")
    assert len(times) > 50 and got.shape == want.shape and np.array_equal(got, want)
    s = ns["solver"]
    assert s.size == info["size"] and s.step == len(times) and np.array_equal(s.h, f.h)
    for d, (i, j, series) in zip(s.recorded_energies, f.probes):
        assert d["energies"] == series
    assert os.path.exists(name + ".txt") and out.getvalue().strip()
