"""GPU: Solver.record_field -- the h series of an on-device probe against the oracle's snapshot
stack, bitwise, with NO snapshot taken on the device.  Written after round 1's GPU budget had run out: it only
combines paths the GPU suite already validates (probes in resident and per-call mode, temporal blocking), and the file
sorts last so that it cannot mask another test under `pytest -x`."""
import contextlib
import io

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("name,n_calls,kw", [("tutorial", 1200, {}), ("waveguide", 900, {}),
                                            ("tutorial", 600, {"resident": False}),
                                            ("waveguide", 500, {"resident": False, "temporal_block": 4})])
def test_record_field_series_equals_the_saved_stack(name, n_calls, kw):
    from fdtd_em_solver_b200 import solver as S
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    times = times[:n_calls]
    with contextlib.redirect_stdout(io.StringIO()):
        s = SC.drive(S, sc, **kw) if kw else SC.drive(S, sc)
        s.end_time = float(times[-1]) + 0.5 * s.dt
        src = s.pulses[0]
        places = [((src.get_row() + 0.5) * s.ds, (src.get_col() + 0.5) * s.ds),        # AT the point source: App. A-Q8
                  (0.4 * s.length_x, 0.6 * s.length_y), (0.0, 0.0)]
        for loc in places:
            s.record_field(loc)
        s.solve(solve_only=True, step_frequency=sc["step_frequency"])
    want = np.array(R.run_saving(f, setup, times, sc["step_frequency"]))
    assert s.h_list == [] and np.array_equal(s.h, f.h)
    for k, loc in enumerate(places):
        dc = s.field_collector(k)
        assert np.array_equal(dc.data, want[:, dc.i_pos, dc.j_pos]), (name, loc)
        assert len(dc.time_vector) == len(want)
    for d, (i, j, series) in zip(s.recorded_energies, f.probes):
        assert d["energies"] == series
