"""CPU, world_size 2 and 3 over gloo: the drop-in Solver in a SHARDED run (one y-slab per rank, shard.py).  The device
engine is replaced by tests/fake_engine.FakeSlabEngine, which steps its slab with the oracle's slab scheme and exchanges
halo rows with the product's own plan (slab.exchange) -- so what is checked is the Solver's sharding logic end to end:
slab assignment, NCCL-id hand-over, the .npy written by all ranks together, solver.h gathered on every rank, probes
taken from the rank that owns the cell, and the update() path; all bitwise against the monolithic oracle."""
import contextlib
import io
import os
import socket

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _quiet(fn, *a, **k):
    with contextlib.redirect_stdout(io.StringIO()):
        return fn(*a, **k)


def _worker(rank, world, port, name, n_calls, out_dir):
    import sys
    sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
    import torch.distributed as dist
    from fake_engine import FakeSlabEngine
    from fdtd_em_solver_b200 import solver as S, analyser as A
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    os.environ["FDTD_SHARDED"] = "1"              # the opt-in an unchanged script gets its sharding from
    S.Engine = FakeSlabEngine
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    times = times[:n_calls]
    k = sc["step_frequency"]

    def driven():
        s = _quiet(SC.drive, S, sc)
        s.end_time = float(times[-1]) + 0.5 * s.dt
        return s

    # ---- a saving run: every rank writes its rows of <name>.npy, rank 0 the .txt -----------------------------
    s = driven()
    src = s.pulses[0]
    places = [((src.get_row() + 0.5) * s.ds, (src.get_col() + 0.5) * s.ds), (0.9 * s.length_x, 0.3 * s.length_y)]
    for loc in places:
        s.record_field(loc)
    path = os.path.join(out_dir, name)
    s.save(path)
    _quiet(s.solve, realtime=False, step_frequency=k)
    want = np.array(R.run_saving(f, setup, times, k))
    assert s._shard_ctx is not None and s._shard_ctx.world == world and len(s._slabs) == world
    assert np.array_equal(np.load(path + ".npy"), want)                          # the file all ranks wrote together
    assert len(s.h_list) == len(want) and all(np.array_equal(a, b) for a, b in zip(s.h_list, want))
    assert np.array_equal(s.h, f.h) and np.array_equal(s.ex, f.ex) and np.array_equal(s.ey, f.ey)   # gathered slabs
    assert s.step == len(times)
    for d, (i, j, series) in zip(s.recorded_energies, f.probes):                 # from the rank that owns the cell
        assert d["energies"] == series
    loader = _quiet(A.FileLoader, path)
    for n, loc in enumerate(places):
        assert np.array_equal(s.field_collector(n).data, loader.create_data_collector(loc).data)
    assert os.path.exists(path + ".txt")

    # ---- solve_only, then a few update() calls on the time axis (engine reused, fields re-zeroed) -------------
    f2, setup2, _, _ = SC.restate(sc)
    s2 = driven()
    _quiet(s2.solve, solve_only=True, step_frequency=k)
    R.run(f2, setup2, times)
    assert np.array_equal(s2.h, f2.h) and np.array_equal(s2.ey, f2.ey)
    s2.reset()
    f3, setup3, _, _ = SC.restate(sc)
    for t in times[:12]:
        R.leapfrog(f3, setup3, t)
        assert np.array_equal(s2.update(t), f3.h)
    with pytest.raises(Exception, match="not available in a sharded run"):
        s2.plot_and_show()
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,name,n_calls", [(2, "waveguide", 260), (3, "tutorial", 230), (2, "mixed", 200)])
def test_sharded_solver_over_gloo_equals_the_monolithic_oracle(world, name, n_calls, tmp_path):
    import torch.multiprocessing as mp
    mp.spawn(_worker, args=(world, _free_port(), name, n_calls, str(tmp_path)), nprocs=world, join=True)


def test_single_process_runs_are_not_sharded(monkeypatch):
    from fdtd_em_solver_b200 import shard
    monkeypatch.delenv("WORLD_SIZE", raising=False)
    monkeypatch.delenv("FDTD_SHARDED", raising=False)
    assert shard.context(None) is None and shard.context(False) is None
    with pytest.raises(RuntimeError):
        shard.context(True)
    monkeypatch.setenv("FDTD_SHARDED", "1")       # the opt-in by environment never fails a single-process run
    assert shard.context(None) is None
    monkeypatch.setenv("WORLD_SIZE", "2")         # ... and without it a torchrun launch stays unsharded (independent runs)
    monkeypatch.delenv("FDTD_SHARDED")
    assert shard.context(None) is None
    assert shard.owner_of_row([(0, 4), (4, 7), (7, 10)], 6) == 1


@pytest.mark.gpu
def test_sharded_solver_under_torchrun_matches_one_gpu():
    """The real thing: the drop-in Solver sharding itself under torchrun (NCCL inside the library) against the same
    scenes on one GPU.  Needs >= 2 GPUs (`gpurun --gpus 2`); skipped on the one-GPU box of the round-end run."""
    import subprocess
    import sys
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % min(n, 4),
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(root, "tests", "sharded_solver_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=600, cwd=root)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SOLVER_SHARDED_OK" in res.stdout
