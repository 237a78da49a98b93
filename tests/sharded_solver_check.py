"""Run under torchrun on 2+ GPUs: the drop-in Solver, launched unchanged, shards itself (shard.py: one y-slab per rank,
halo rows over NCCL inside libfdtd2d) and must leave the same .npy / fields / probe series as the same script on ONE GPU.

    torchrun --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 tests/sharded_solver_check.py

The one-GPU run (Solver(..., sharded=False) on every rank's own GPU) is the reference here: it is the path the GPU
parity suite pins bitwise to the oracle.  Test infrastructure: oracle.scenarios supplies the scene descriptions."""
import contextlib
import io
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np

from fdtd_em_solver_b200 import solver as S
from oracle import scenarios as SC

rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
tmp = os.environ.get("FDTD_SHARD_TMP") or os.path.join(tempfile.gettempdir(), "fdtd_sharded_check")
os.makedirs(tmp, exist_ok=True)

for name, n_calls, kw in [("waveguide", 900, {}), ("tutorial", 1200, {}), ("doubleslit", 1500, {}),
                          ("lens", 700, {"temporal_block": 4}), ("mixed", 600, {"dtype": "float32"})]:
    sc = SC.spec(name)
    runs = []
    for sharded in (True, False):
        with contextlib.redirect_stdout(io.StringIO()):
            s = SC.drive(S, sc, sharded=sharded, device=int(os.environ.get("LOCAL_RANK", "0")), **kw)
            s.end_time = (n_calls - 0.5) * s.dt
            k = s.record_field((0.5 * s.length_x, 0.25 * s.length_y))
            path = os.path.join(tmp, "%s_%s_r%d" % (name, "sharded" if sharded else "single", 0 if sharded else rank))
            s.save(path)
            t0 = time.perf_counter()
            s.solve(realtime=False, step_frequency=sc["step_frequency"])
            dt = time.perf_counter() - t0
        assert (s._shard_ctx is not None) == sharded
        runs.append((np.load(path + ".npy"), s.h, s.ex, s.ey, s.field_collector(k).data,
                     [d["energies"] for d in s.recorded_energies], dt))
    a, b = runs
    ok = all(np.array_equal(x, y) for x, y in zip(a[:5], b[:5])) and a[5] == b[5]
    if rank == 0:
        print("%s %s x%d ranks, %d calls: %s  (sharded %.2f s, one GPU %.2f s)" % (
            name, kw, world, n_calls, "bit-identical" if ok else "MISMATCH", a[6], b[6]), flush=True)
    assert ok
if rank == 0:
    print("SOLVER_SHARDED_OK", flush=True)
import torch.distributed as dist
dist.barrier()
dist.destroy_process_group()
