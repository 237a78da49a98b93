"""Per-call time of BASELINE.json configs 1-4 through the drop-in Solver API: per-call launches (resident=False) vs
resident mode (one cooperative launch per snapshot interval / per run).  GPU only; prints one JSON line per config."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import scenarios as SC                     # scenario specs only (no oracle stepping here)
from fdtd_em_solver_b200 import solver as S

import torch
torch.zeros(1, device="cuda"); torch.cuda.synchronize()
from fdtd_em_solver_b200 import _lib
_lib.load()

names = sys.argv[1:] or ["waveguide", "lens", "tutorial", "doubleslit"]
for name in names:
    sc = SC.spec(name)
    row = dict(config=name)
    finals = {}
    for resident in (False, True):
        with contextlib.redirect_stdout(io.StringIO()):
            s = SC.drive(S, sc, resident=resident)
            t0 = time.perf_counter()
            s.solve(solve_only=True)
            dt = time.perf_counter() - t0
            st = s._engine.stats()
        key = "resident" if resident else "per_call"
        row.update({key + "_s": round(dt, 4), key + "_us_per_call": round(1e6 * dt / s.step, 2),
                    key + "_device_ms": round(st["last_run_ms"], 3), key + "_launches": st["kernel_launches"]})
        row.update(N=s.size, calls=s.step)
        finals[resident] = (s.h.copy(), s.ex.copy(), s.ey.copy())
    row["bit_identical"] = all(np.array_equal(a, b) for a, b in zip(finals[False], finals[True]))
    if name != "doubleslit":                           # with the snapshot stack written (.npy + .txt), resident mode
        with contextlib.redirect_stdout(io.StringIO()):
            s = SC.drive(S, sc, resident=True)
            s.save("/tmp/_res_" + name)
            t0 = time.perf_counter()
            s.solve(realtime=False, step_frequency=sc["step_frequency"])
            row["resident_save_s"] = round(time.perf_counter() - t0, 4)
    print(json.dumps(row), flush=True)
