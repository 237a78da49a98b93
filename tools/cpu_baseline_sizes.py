"""SURVEY.md 8(d) "CPU baseline beside it": the reference's numpy step (oracle port, bit-identical) and the threaded C
restatement, timed at several grid sizes with the C5 material/source recipe on THIS host.  Reported only.

    python tools/cpu_baseline_sizes.py [--sizes 1024 2048 4096] [--budget 8] > profiles/<name>.json
"""
import argparse
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import c_oracle as CO, restated as R              # noqa: E402  (tool: checker-side code only)
from fdtd_em_solver_b200 import synthetic as SY               # noqa: E402


def timed(step, rows, cols, budget):
    epc, muc = R.synthetic_coefficients(rows, cols, SY.DT, SY.DS, seed=0)
    n_calls = 4096
    pulses = [R.PulseSpec("table", p.direction, p.row, p.col, p.start, p.end, p.table)
              for p in SY.pulses(rows, cols, n_calls)]
    setup = R.Setup(eps_coef=epc, mu_coef=muc, courant=(SY.C0 * SY.DT / SY.DS), pulses=pulses)
    f = R.Fields.zeros(rows, cols)
    t = SY.times(n_calls)
    step(f, setup, t[0])
    n, t0 = 0, time.perf_counter()
    while n < n_calls - 1 and (n < 3 or time.perf_counter() - t0 < budget):
        step(f, setup, t[n + 1])
        n += 1
    dt = time.perf_counter() - t0
    return {"steps": n, "seconds": round(dt, 3), "ms_per_step": 1e3 * dt / n, "mcell_per_s": rows * cols * n / dt / 1e6}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sizes", type=int, nargs="+", default=[1024, 2048, 4096])
    ap.add_argument("--budget", type=float, default=8.0)
    a = ap.parse_args()
    out = {"host_cores": os.cpu_count(), "recipe": "C5 synthetic coefficients + sources, fp64", "sizes": {}}
    for n in a.sizes:
        row = {"numpy_port_1_core": timed(R.leapfrog, n, n, a.budget)}
        if CO.available():
            CO.set_threads(1)
            row["c_port_1_thread"] = timed(CO.leapfrog, n, n, a.budget)
            CO.set_threads(os.cpu_count())
            row["c_port_all_threads"] = timed(CO.leapfrog, n, n, a.budget)
        out["sizes"][str(n)] = row
    print(json.dumps(out, indent=1))


if __name__ == "__main__":
    main()
