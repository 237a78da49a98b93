"""y-slab decomposition: host logic + halo protocol on CPU (gloo, world_size 2 and 3, oracle compute)
and the slab kernels on the GPU (several slab contexts on one device, halo rows copied between them)."""
import os
import socket
import subprocess
import sys

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC, slabs as OS
from fdtd_em_solver_b200 import slab as SL

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_split_rows_and_plan_symmetry():
    assert SL.split_rows(10, 3) == [(0, 4), (4, 7), (7, 10)]
    assert SL.split_rows(65536, 8)[3] == (24576, 32768)
    with pytest.raises(ValueError):
        SL.split_rows(5, 3)
    for world in (2, 3, 8):
        plans = [SL.plan_for(r, world, 64, 40) for r in range(world)]
        for p in plans:
            assert p.local_rows == (p.row_end - p.row_begin) + p.halo_top + 1
            for m in p.sends():          # every send is matched, in order, by a recv on the peer
                peer = plans[m.peer]
                mine = [x for x in p.sends() if x.peer == m.peer]
                theirs = [x for x in peer.recvs() if x.peer == p.rank]
                assert [(x.field, x.row, x.count) for x in mine] == [(x.field, x.row, x.count) for x in theirs]
        assert plans[0].bytes_per_step(8) == (3 * 40 + 1) * 8
        assert plans[1].bytes_per_step(8) == ((3 * 40 + 1) * 8 if world > 2 else 0) + 40 * 8


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _case(rows=37, cols=29, seed=2, n_steps=14):
    return SC.random_case(rows, cols, seed, walls=(False, False, True, False), n_steps=n_steps)


def _gloo_worker(rank, world, port, rows, cols, seed, n_steps, out_dir):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    f, setup, times = SC.random_case(rows, cols, seed, walls=(False, False, True, False), n_steps=n_steps)
    plan = SL.plan_for(rank, world, rows, cols)
    fl, off = OS.make_slab(f, plan.row_begin, plan.row_end)
    assert off == plan.row_off and fl.ex.shape[0] == plan.local_rows
    sl = OS.local_setup(setup, off, fl.h.shape[0], rows)
    for t in times:
        OS.leapfrog_slab(fl, sl, t, rows, off)
        arrays = {"h": torch.from_numpy(fl.h), "ex": torch.from_numpy(fl.ex), "ey": torch.from_numpy(fl.ey)}
        SL.exchange(plan, arrays, dist)
    ht = plan.halo_top
    n = plan.row_end - plan.row_begin
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), h=fl.h[ht:ht + n],
             ex=fl.ex[ht:ht + n + (1 if plan.row_end == rows else 0)], ey=fl.ey[ht:ht + n])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_protocol_over_gloo_matches_monolithic_oracle(world, tmp_path):
    import torch.multiprocessing as mp
    rows, cols, seed, n_steps = 37, 29, 2, 14
    mp.spawn(_gloo_worker, args=(world, _free_port(), rows, cols, seed, n_steps, str(tmp_path)), nprocs=world, join=True)
    f, setup, times = _case(rows, cols, seed, n_steps)
    for t in times:
        R.leapfrog(f, setup, t)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(world)]
    assert np.array_equal(np.concatenate([p["h"] for p in parts]), f.h)
    assert np.array_equal(np.concatenate([p["ex"] for p in parts]), f.ex)
    assert np.array_equal(np.concatenate([p["ey"] for p in parts]), f.ey)


# ------------------------------------------------------------------------------------------ GPU
def _slab_engines(f, setup, times, world, **kw):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = f.h.shape
    S = setup.courant
    engines = []
    for r in range(world):
        b, e_ = SL.split_rows(rows, world)[r]
        e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls,
                   row_begin=b, row_end=e_, **kw)
        e.upload_coefficients(setup.eps_coef, setup.mu_coef)
        e.set_sources(setup.pulses, times)
        e.set_reflectors(setup.reflectors)
        e.set_fields(f.h, f.ex, f.ey)
        engines.append(e)
    return engines


def _copy_halos(engines):
    """What fdtd_halo_exchange does over NCCL, done here with device-to-device row copies."""
    world = len(engines)
    for r, e in enumerate(engines):
        bufs = dict(zip(("h", "ex", "ey"), e.generation_buffers()))
        plan = SL.plan_for(r, world, e.rows, e.cols)
        for m in plan.sends():
            peer = engines[m.peer]
            pp = SL.plan_for(m.peer, world, e.rows, e.cols)
            pb = dict(zip(("h", "ex", "ey"), peer.generation_buffers()))
            pb[m.field][m.row - pp.row_off, :m.count].copy_(bufs[m.field][m.row - plan.row_off, :m.count])
    import torch
    torch.cuda.synchronize()


@pytest.mark.gpu
@pytest.mark.parametrize("world,kernel", [(2, "stream"), (3, "stream"), (4, "exact"), (5, "stream")])
def test_slab_kernels_on_one_gpu_match_oracle(world, kernel):
    f, setup, times = SC.random_case(41, 150, 7, walls=(False, False, True, False), n_steps=9)
    engines = _slab_engines(f, setup, times, world, kernel=kernel, tile_rows=4)
    for n, t in enumerate(times):
        R.leapfrog(f, setup, t)
        for e in engines:
            e.step(n)
        for e in engines:
            e.sync()
        _copy_halos(engines)
    h, ex, ey = np.zeros_like(f.h), np.zeros_like(f.ex), np.zeros_like(f.ey)
    for e in engines:
        e.get_fields(out={"h": h, "ex": ex, "ey": ey})
    assert np.array_equal(h, f.h) and np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
    for e in engines:
        e.close()


@pytest.mark.gpu
def test_nccl_slabs_under_torchrun_match_single_gpu():
    """The real multi-process path: one rank per GPU, halo rows over NCCL inside the C library."""
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs >= 2 GPUs (run with gpurun --gpus 2)")
    world = 2 if n < 4 else 4
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=%d" % world,
           "--master-addr", "127.0.0.1", "--master-port", str(_free_port()), os.path.join(ROOT, "tools", "slab_check.py")]
    res = subprocess.run(cmd, capture_output=True, text=True, timeout=900, cwd=ROOT)
    assert res.returncode == 0, res.stdout[-3000:] + res.stderr[-3000:]
    assert "SLAB_CHECK_OK" in res.stdout


# ------------------------------------------------------------------------------------------ deep halos (CPU, gloo)
def _deep_worker(rank, world, port, rows, cols, seed, K, n_steps, out_dir):
    import torch
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    f, setup, times = SC.random_case(rows, cols, seed, walls=(False, False, True, False), n_steps=n_steps)
    D = K + 1
    plan = SL.deep_plan_for(rank, world, rows, D)
    off, b, e = plan.row_off, plan.row_begin, plan.row_end
    hi = e + plan.halo_bot
    fl = R.Fields(f.h[off:hi].copy(), f.ex[off:hi + 1].copy(), f.ey[off:hi].copy(), step=f.step)
    assert fl.ex.shape[0] == plan.local_rows
    sl = OS.local_setup(setup, off, fl.h.shape[0], rows)
    n = 0
    while n < n_steps:
        g = K if n_steps - n >= K else 1                     # whole groups without any exchange inside, then single steps
        for t in times[n:n + g]:
            fl.step = n
            OS.leapfrog_slab(fl, sl, t, rows, off)
            n += 1
        arrays = {"h": torch.from_numpy(fl.h), "ex": torch.from_numpy(fl.ex), "ey": torch.from_numpy(fl.ey)}
        SL.exchange_deep(plan, arrays, dist)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), h=fl.h[b - off:e - off],
             ex=fl.ex[b - off:e - off + (1 if e == rows else 0)], ey=fl.ey[b - off:e - off])
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("world,K", [(2, 3), (3, 2), (3, 6)])
def test_deep_halo_scheme_over_gloo_matches_monolithic_oracle(world, K, tmp_path):
    """K+1 rows of each neighbour, ONE exchange per group of K steps (and one per leftover single step): the owned rows
    stay bit-identical to the monolithic run -- what the stale edge of the stored rows computes moves one row per step
    and so never reaches an owned row inside a group."""
    import torch.multiprocessing as mp
    rows, cols, seed, n_steps = 61, 29, 2, 3 * K + 2
    mp.spawn(_deep_worker, args=(world, _free_port(), rows, cols, seed, K, n_steps, str(tmp_path)), nprocs=world, join=True)
    f, setup, times = SC.random_case(rows, cols, seed, walls=(False, False, True, False), n_steps=n_steps)
    for t in times:
        R.leapfrog(f, setup, t)
    parts = [np.load(str(tmp_path / ("rank%d.npz" % r))) for r in range(world)]
    assert np.array_equal(np.concatenate([p["h"] for p in parts]), f.h)
    assert np.array_equal(np.concatenate([p["ex"] for p in parts]), f.ex)
    assert np.array_equal(np.concatenate([p["ey"] for p in parts]), f.ey)


def test_deep_plan_matches_the_library_layout():
    from fdtd_em_solver_b200 import _lib as L
    import ctypes as C
    lib = L.load()
    for world, rows, depth in ((2, 100, 7), (3, 61, 4), (8, 65536, 7)):
        for r in range(world):
            p = SL.deep_plan_for(r, world, rows, depth)
            cfg = L.Config()
            cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end, cfg.dtype = rows, 40, p.row_begin, p.row_end, L.F64
            assert p.local_rows == lib.fdtd_local_rows_halo(C.byref(cfg), depth)
            for peer, send_row, recv_row in p.blocks():
                q = SL.deep_plan_for(peer, world, rows, depth)
                assert (r, recv_row, send_row) in [(x[0], x[1], x[2]) for x in q.blocks()]     # my receive is its send
