"""Tuning sweep of the streaming kernel on the C5 slab (run on the GPU box)."""
import argparse, json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from fdtd_em_solver_b200.engine import Engine, C0
from fdtd_em_solver_b200 import synthetic

ap = argparse.ArgumentParser()
ap.add_argument("--rows", type=int, default=8192)
ap.add_argument("--cols", type=int, default=65536)
ap.add_argument("--dtype", default="float64")
ap.add_argument("--steps", type=int, default=20)
ap.add_argument("--configs", default="")
ap.add_argument("--temporal", type=int, default=1)
a = ap.parse_args()
e, times = synthetic.make_engine(a.rows, a.cols, dtype=a.dtype, n_calls=4096, temporal=a.temporal)
bytes_per_cell = 64 if a.dtype == "float64" else 32
va = 2 if a.dtype == "float64" else 4
cfgs = [(v, tr, w) for v in (va, 2 * va) for tr in (32, 64, 128, 256) for w in (2, 4, 8)]
if a.configs:
    cfgs = [tuple(int(x) for x in c.split(",")) for c in a.configs.split(";")]
n = 0
for v, tr, w in cfgs:
    e.set_tuning(tr, v, w)
    e.run(n, 3); n += 3
    e.run(n, a.steps); n += a.steps
    ms = e.stats()["last_run_ms"] / a.steps
    g = a.rows * a.cols / ms / 1e6
    print(json.dumps(dict(temporal=a.temporal, vec=v, tile_rows=tr, warps=w, ms_per_step=round(ms, 4), gcells=round(g, 2),
                          gbs=round(g * bytes_per_cell, 1))), flush=True)
