"""Run under torchrun: every rank steps its y-slab (halo rows over NCCL inside libfdtd2d); rank 0 also
steps the whole grid on its own GPU and the gathered slabs must be bit-identical.

    torchrun --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 tools/slab_check.py [--big]

Small cases: seeded all-feature scenes (every pulse kind, reflectors, a reflecting wall) in every launch mode -- single
steps, the library's defaults under temporal blocking (deep halos of K+1 rows, ONE exchange per group, stepK_lean +
stepK_edge) and the classic one-row halos (exchange every step; masked pass chain through scratch generations).
--big adds the sizes SURVEY.md 8(d) names: 4096 x 4096 with all features, and the C5 recipe on 16 384 x 65 536
(1 GPU == N GPUs, compared on the device: the slabs travel to rank 0 over NCCL)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from fdtd_em_solver_b200.engine import Engine, comm_unique_id
from fdtd_em_solver_b200 import slab as SL, synthetic as SY
from oracle import scenarios as SC           # only to build a seeded random case (inputs), not as a compute path

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
BIG = "--big" in sys.argv[1:]


def connect(eng):
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    eng.comm_init(rank, world, ident.cpu().numpy().tobytes())


CASES = [(97, 300, 12, dict(tile_rows=4)), (1500, 4100, 25, dict()),
         # the defaults under temporal blocking: deep halos, stepK_lean + stepK_edge
         (97, 300, 13, dict(tile_rows=4, temporal=2)), (1500, 4100, 25, dict(temporal=2)), (1500, 4100, 26, dict(temporal=4)),
         (2000, 9000, 19, dict(temporal=3, block_warps=2)), (2000, 9000, 29, dict(temporal=6)),
         (2000, 9000, 27, dict(temporal=6, tile_rows=16)), (1500, 4100, 35, dict(temporal=8)),
         (1500, 4100, 35, dict(temporal=8, kstep_variant=0)), (3000, 9000, 43, dict(temporal=8, tile_rows=128)),
         # the classic layout: one halo row, exchanged every step (and after every masked pass of a group)
         (1500, 4100, 25, dict(temporal=2, halo_depth=1, frame_mode=0)), (2000, 9000, 27, dict(temporal=6, tile_rows=16, halo_depth=1, frame_mode=0)),
         (1500, 4100, 33, dict(temporal=8, halo_depth=1, frame_mode=0)),
         # a TF/SF "right" line at column 0 (kept in this case only): stepK_edge does not take it -> call by call
         (1500, 4100, 20, dict(temporal=4, keep_wrapping_line=True))]
if BIG:
    CASES.append((4096, 4096, 41, dict(temporal=8)))

for rows, cols, n_steps, kw in CASES:
    kw = dict(kw)
    f, setup, times = SC.random_case(rows, cols, 3, walls=(False, False, False, True), n_steps=n_steps)
    if not kw.pop("keep_wrapping_line", False):
        setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    b, e_ = SL.split_rows(rows, world)[rank]
    S = setup.courant
    eng = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, row_begin=b,
                 row_end=e_, device=local, **kw)
    eng.upload_coefficients(setup.eps_coef, setup.mu_coef)
    eng.set_sources(setup.pulses, times)
    eng.set_reflectors(setup.reflectors)
    eng.set_fields(f.h, f.ex, f.ey)
    connect(eng)
    eng.run(0, n_steps)
    h, ex, ey = eng.get_fields()
    parts = [None] * world
    dist.all_gather_object(parts, (b, e_, h[b:e_], ex[b:e_ + (1 if e_ == rows else 0)], ey[b:e_]))
    if rank == 0:
        whole = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, device=local,
                       **{k: v for k, v in kw.items() if k != "halo_depth"})
        whole.upload_coefficients(setup.eps_coef, setup.mu_coef)
        whole.set_sources(setup.pulses, times)
        whole.set_reflectors(setup.reflectors)
        whole.set_fields(f.h, f.ex, f.ey)
        whole.run(0, n_steps)
        H, EX, EY = whole.get_fields()
        gh = np.concatenate([p[2] for p in parts]); gex = np.concatenate([p[3] for p in parts])
        gey = np.concatenate([p[4] for p in parts])
        ok = np.array_equal(gh, H) and np.array_equal(gex, EX) and np.array_equal(gey, EY)
        print("grid %dx%d world %d steps %d %s halo_depth %d launches %d: %s  halo bytes/step/rank0 %d" % (
            rows, cols, world, n_steps, kw, eng.halo_depth, eng.stats()["kernel_launches"],
            "bit-identical" if ok else "MISMATCH", eng.stats()["halo_bytes_sent"] // n_steps), flush=True)
        assert ok
        whole.close()
    eng.close()
    dist.barrier()

if BIG:
    # "1-GPU == G-GPU bitwise on 16 384 x 65 536" -- the C5 recipe, defaults (K = 8), compared on the device
    rows, cols, n_steps = 16384, 65536, 27
    per = rows // world
    edges = tuple(r * per for r in range(world))
    eng, _ = SY.make_engine(rows, cols, n_calls=64, row_begin=rank * per, row_end=(rank + 1) * per, slab_edges=edges,
                            device=local, temporal=8)
    connect(eng)
    eng.run(0, n_steps)
    mine = [t[eng.row_begin - eng.row_off:eng.row_end - eng.row_off] for t in eng.generation_buffers()]
    whole = None
    if rank == 0:
        whole, _ = SY.make_engine(rows, cols, n_calls=64, slab_edges=edges, device=local, temporal=8)
        whole.run(0, n_steps)
    ok = True
    for r in range(world):
        for k in range(3):
            if rank == 0:
                ref = whole.generation_buffers()[k][r * per:(r + 1) * per]
                if r == 0:
                    got = mine[k]
                else:
                    got = torch.empty_like(ref)
                    dist.recv(got, src=r)
                ok = ok and bool(torch.equal(got, ref))
                nonzero = float(ref.abs().max()) > 0 if k == 0 and r == 0 else True
                ok = ok and nonzero
            elif rank == r:
                dist.send(mine[k].contiguous(), dst=0)
    if rank == 0:
        print("C5 recipe %dx%d world %d steps %d (defaults: K=8, halo_depth %d): 1 GPU == %d GPUs %s" % (
            rows, cols, world, n_steps, eng.halo_depth, world, "bit-identical" if ok else "MISMATCH"), flush=True)
        assert ok
        whole.close()
    eng.close()
    dist.barrier()
if rank == 0:
    print("SLAB_CHECK_OK", flush=True)
dist.destroy_process_group()
