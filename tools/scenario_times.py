"""Wall-clock of BASELINE.json configs 1-4 through the drop-in Solver API on the GPU vs the numpy oracle port
on the host (same box, same run).  These grids (139..319 cells per side) are launch/latency-bound on a GPU:
parity configs, not roofline configs."""
import contextlib, io, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from oracle import restated as R, scenarios as SC
from fdtd_em_solver_b200 import solver as S

import torch
torch.zeros(1, device="cuda"); torch.cuda.synchronize()          # CUDA context creation is not part of any config's time
from fdtd_em_solver_b200 import _lib
_lib.load()

for name in ("tutorial", "lens", "waveguide", "doubleslit"):
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    n_cpu = min(len(times), 3000)
    t0 = time.perf_counter()
    for t in times[:n_cpu]:
        R.leapfrog(f, setup, t)
    cpu_s = (time.perf_counter() - t0) * len(times) / n_cpu
    with contextlib.redirect_stdout(io.StringIO()):
        s = SC.drive(S, sc)
        s.save("/tmp/_scn_" + name)
        t0 = time.perf_counter()
        s.solve(realtime=False, step_frequency=sc["step_frequency"])
        gpu_save_s = time.perf_counter() - t0
        s2 = SC.drive(S, sc)
        t0 = time.perf_counter()
        s2.solve(solve_only=True)
        gpu_s = time.perf_counter() - t0
    n = info["size"]
    print(json.dumps(dict(config=name, N=n, calls=len(times), numpy_s=round(cpu_s, 2),
                          gpu_solve_only_s=round(gpu_s, 3), gpu_save_s=round(gpu_save_s, 3),
                          gpu_us_per_step=round(1e6 * gpu_s / len(times), 2),
                          gpu_mcells=round(n * n * len(times) / gpu_s / 1e6, 1),
                          numpy_mcells=round(n * n * len(times) / cpu_s / 1e6, 1))), flush=True)
