"""Run under torchrun: every rank steps its y-slab (halo rows over NCCL inside libfdtd2d); rank 0 also
steps the whole grid on its own GPU and the gathered slabs must be bit-identical (fields and checksums)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import torch
import torch.distributed as dist

from fdtd_em_solver_b200.engine import Engine, comm_unique_id
from fdtd_em_solver_b200 import slab as SL
from oracle import scenarios as SC           # only to build a seeded random case (inputs), not as a compute path

rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))

# frame modes 1/2 (frame_window + interface pass chain, window.cuh) join the list with FDTD_SLAB_CHECK_FRAME=1
FRAME_CASES = [(1500, 4100, 26, dict(temporal=4, frame_mode=1)), (2000, 9000, 27, dict(temporal=6, tile_rows=16, frame_mode=1)),
               (2000, 9000, 27, dict(temporal=6, tile_rows=16, frame_mode=2)), (1500, 4100, 33, dict(temporal=8, frame_mode=2)),
               (97, 300, 13, dict(tile_rows=4, temporal=2, frame_mode=1))] if os.environ.get("FDTD_SLAB_CHECK_FRAME") == "1" else []

# deep halos (one exchange per K-step group, fdtd_set_halo_depth) join with FDTD_SLAB_CHECK_DEEP=1
DEEP_CASES = [(1500, 4100, 26, dict(temporal=4, halo_depth=5)), (2000, 9000, 27, dict(temporal=6, tile_rows=16, halo_depth=7)),
              (2000, 9000, 29, dict(temporal=6, halo_depth=7, frame_mode=1)), (1500, 4100, 35, dict(temporal=8, halo_depth=9)),
              (97, 300, 13, dict(tile_rows=4, temporal=2, halo_depth=3)),
              # frame mode 3: plain K-step tiles + stepK_edge tiles, nothing else
              (1500, 4100, 26, dict(temporal=4, halo_depth=5, frame_mode=3)), (2000, 9000, 29, dict(temporal=6, halo_depth=7, frame_mode=3)),
              (1500, 4100, 35, dict(temporal=8, halo_depth=9, frame_mode=3, kstep_variant=1)),
              (97, 300, 13, dict(tile_rows=4, temporal=2, halo_depth=3, frame_mode=3))] if os.environ.get("FDTD_SLAB_CHECK_DEEP") == "1" else []

for rows, cols, n_steps, kw in DEEP_CASES + FRAME_CASES + [(97, 300, 12, dict(tile_rows=4)), (1500, 4100, 25, dict()),
                                (97, 300, 13, dict(tile_rows=4, temporal=2)), (1500, 4100, 25, dict(temporal=2)),
                                (2000, 9000, 16, dict(temporal=2, block_warps=2)),
                                (1500, 4100, 26, dict(temporal=4)), (2000, 9000, 19, dict(temporal=3, block_warps=2)),
                                (2000, 9000, 27, dict(temporal=6, tile_rows=16)), (1500, 4100, 33, dict(temporal=8))]:
    f, setup, times = SC.random_case(rows, cols, 3, walls=(False, False, False, True), n_steps=n_steps)
    if kw.get("frame_mode") or kw.get("halo_depth"):         # a TF/SF line at column 0 would keep the whole frame on the masked passes
        setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    b, e_ = SL.split_rows(rows, world)[rank]
    S = setup.courant
    eng = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, row_begin=b,
                 row_end=e_, device=local, **kw)
    eng.upload_coefficients(setup.eps_coef, setup.mu_coef)
    eng.set_sources(setup.pulses, times)
    eng.set_reflectors(setup.reflectors)
    eng.set_fields(f.h, f.ex, f.ey)
    ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
    if rank == 0:
        ident.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
    dist.broadcast(ident, 0)
    eng.comm_init(rank, world, ident.cpu().numpy().tobytes())
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
        print("grid %dx%d world %d steps %d %s: %s  halo bytes/step/rank0 %d" % (
            rows, cols, world, n_steps, kw, "bit-identical" if ok else "MISMATCH",
            eng.stats()["halo_bytes_sent"] // n_steps), flush=True)
        assert ok
        whole.close()
    eng.close()
    dist.barrier()
if rank == 0:
    print("SLAB_CHECK_OK", flush=True)
dist.destroy_process_group()
