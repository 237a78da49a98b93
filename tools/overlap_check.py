"""Temporal blocking K=6 on the C5 slab (8192 x 65536 fp64): last frame pass queued behind the K-step kernel (default)
vs on the edge stream next to it (FDTD_OVERLAP_LAST=1).  Prints parity (bitwise, on the device) and Gcell/s of both."""
import json, os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fdtd_em_solver_b200 import synthetic as SY

rows, cols = int(os.environ.get("ROWS", 8192)), int(os.environ.get("COLS", 65536))
steps, K = int(os.environ.get("STEPS", 120)), int(os.environ.get("TEMPORAL", 6))
out = {}
engines = {}
for overlap in ("0", "1"):
    os.environ["FDTD_OVERLAP_LAST"] = overlap          # read by fdtd_create
    e, _ = SY.make_engine(rows, cols, n_calls=steps + 64, temporal=K)
    e.run(0, 19)                                        # parity leg: 3 groups + 1 single step (also the warm-up)
    engines[overlap] = e
fa, fb = engines["0"].generation_buffers(), engines["1"].generation_buffers()
out["bit_identical"] = all(bool(torch.equal(x, y)) for x, y in zip(fa, fb))
out["nontrivial"] = float(fa[0].abs().max()) > 0
for overlap, e in engines.items():
    best = None
    for rep in range(2):
        e.run(19 + rep * 0, steps)                      # same calls twice: the time does not depend on the values
        ms = e.stats()["last_run_ms"]
        best = ms if best is None else min(best, ms)
    out["overlap_%s_ms_per_step" % overlap] = round(best / steps, 4)
    out["overlap_%s_gcells" % overlap] = round(rows * cols * steps / best / 1e6, 1)
print(json.dumps(out), flush=True)
