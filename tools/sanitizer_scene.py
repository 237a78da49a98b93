"""One small scene through every kernel family of libfdtd2d.so, each checked bitwise against the oracle.  Meant to run
under `compute-sanitizer --tool memcheck|racecheck|synccheck|initcheck` (the shared-memory kernels --
the cp.async rings of stepK_stream / stepK_lean, the EdgeWallState / pulse records of stepK_edge, the tiles of
step_resident -- have no race or out-of-bounds access); without a sanitizer it is a 2-second smoke run.

    compute-sanitizer --tool racecheck python tools/sanitizer_scene.py [family ...]
"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from oracle import restated as R, scenarios as SC  # noqa: E402
from fdtd_em_solver_b200.engine import Engine  # noqa: E402

WALLS = (False, True, False, False)


def run(family, dtype="float64"):
    rows, cols, steps = 150, 400, 17
    kw = dict(step_stream=dict(kernel="stream", tile_rows=8, block_warps=2, resident=False),
              step_exact=dict(kernel="exact", resident=False),
              stepK_stream=dict(kernel="stream", tile_rows=16, temporal=6, kstep_variant=0, frame_mode=0, resident=False),
              stepK_lean_edge=dict(kernel="stream", tile_rows=16, temporal=8, kstep_variant=1, frame_mode=3, resident=False),
              stepK_lean_edge_k3=dict(kernel="stream", tile_rows=24, temporal=3, kstep_variant=1, frame_mode=3, resident=False),
              step_resident=dict(kernel="stream", resident=True),
              split_storage=dict(kernel="stream"))[family]
    f, setup, times = SC.random_case(rows, cols, 5, walls=WALLS, n_steps=steps)
    cells = [(0, 0), (rows // 2, cols // 2), (rows - 1, cols - 1)]
    e = Engine(rows, cols, courant=setup.courant, one_minus_courant=1 - setup.courant, reflect=setup.reflect_walls,
               dtype="float32x" if family == "split_storage" else dtype, **kw)
    e.upload_coefficients(setup.eps_coef, setup.mu_coef)
    e.set_sources(setup.pulses, times)
    e.set_reflectors(setup.reflectors)
    e.set_fields(f.h, f.ex, f.ey)
    e.set_probes(cells)
    stack = e.run(0, steps, snapshot_every=4, total_calls=steps)
    h, ex, ey = e.get_fields()
    launches = e.stats()["kernel_launches"]
    e.close()
    if family == "split_storage":
        from oracle import split_storage as SS
        want = np.array(SS.run_saving(f, setup, times, 4))
        ok = np.array_equal(stack, want) and np.array_equal(h, f.h) and np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
        dtype = "float32x"
    elif dtype == "float64":
        want = np.array(R.run_saving(f, setup, times, 4))
        ok = np.array_equal(stack, want) and np.array_equal(h, f.h) and np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
    else:
        R.run(f, setup, times)
        ok = np.linalg.norm(h - f.h) <= 1e-5 * np.linalg.norm(f.h)      # fp32: 17 steps stay far inside the gate
    print("%-20s %-8s launches %3d  %s" % (family, dtype, launches, "bit-exact" if ok and dtype != "float32" else
                                            "within 1e-5" if ok else "MISMATCH"), flush=True)
    return ok


if __name__ == "__main__":
    fams = sys.argv[1:] or ["step_stream", "step_exact", "stepK_stream", "stepK_lean_edge", "stepK_lean_edge_k3", "step_resident"]
    good = all([run(fam, dt) for fam in fams for dt in ("float64", "float32")] + ([run("split_storage")] if not sys.argv[1:] else []))
    sys.exit(0 if good else 1)
