"""Per-snapshot relative-L2 error of the fp32 path against the fp64 path (which is bit-identical to the
reference) on BASELINE configs 1-4: worst snapshot, first snapshot over the 1e-5 gate, probe-series error."""
import contextlib, io, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import scenarios as SC
from fdtd_em_solver_b200 import solver as S

for name in ("tutorial", "lens", "waveguide", "doubleslit"):
    sc = SC.spec(name)
    stacks = {}
    for dtype in ("float64", "float32"):
        with contextlib.redirect_stdout(io.StringIO()):
            s = SC.drive(S, sc, dtype=dtype)
            s._coefficients()
            s.time_vector = np.arange(0, s.end_time, s.dt)
            if name == "doubleslit":
                s.time_vector = s.time_vector[:12000]
            stacks[dtype] = s._run_device(s.time_vector, 5)
    a, b = stacks["float64"], stacks["float32"]
    den = np.sqrt((a ** 2).sum(axis=(1, 2)))
    num = np.sqrt(((a - b) ** 2).sum(axis=(1, 2)))
    ok = den > 0
    rel = np.zeros_like(den); rel[ok] = num[ok] / den[ok]
    over = np.nonzero(rel > 1e-5)[0]
    i, j = sc["probe_cells"][0]
    pa, pb = a[:, i, j], b[:, i, j]
    print(json.dumps(dict(config=name, snapshots=len(a), worst_rel_l2=float(rel.max()), worst_at_step=int(5 * (rel.argmax() + 1)),
                          first_step_over_1e5=(int(5 * (over[0] + 1)) if len(over) else None),
                          snapshots_within_1e5=int((rel <= 1e-5).sum()),
                          probe_series_rel_l2=float(np.linalg.norm(pa - pb) / max(np.linalg.norm(pa), 1e-300)))), flush=True)
