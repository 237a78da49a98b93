"""Shared by the example scripts: make the package importable from a checkout and parse the two common switches."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)


def options(description, default_name):
    ap = argparse.ArgumentParser(description=description)
    ap.add_argument("--name", default=default_name, help="basename of the .npy / .txt the run writes")
    ap.add_argument("--show", action="store_true", help="realtime matplotlib view instead of saving (needs a display)")
    ap.add_argument("--time-scale", type=float, default=1.0, help="shorten the simulated time (1.0 = the reference's)")
    ap.add_argument("--dtype", default="float64", choices=["float64", "float32", "float32x"],
                    help="float32x: fp32 hi + bf16 lo field storage with float64 arithmetic (stays within 1e-5 of float64)")
    return ap.parse_args()


def run(solver, opt, step_frequency=5):
    """save + headless solve (the reference scripts' `solver.save(...)`, `solver.solve(realtime=False)`), or the
    realtime view with --show.  Under `torchrun --nproc-per-node N` the solver shards itself into y-slabs."""
    if opt.show:
        solver.solve(realtime=True, step_frequency=step_frequency)
        return None
    solver.save(opt.name)
    solver.solve(realtime=False, step_frequency=step_frequency)
    from fdtd_em_solver_b200.analyser import FileLoader
    return FileLoader(opt.name)
