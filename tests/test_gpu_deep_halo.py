"""GPU: deep halos (fdtd_set_halo_depth, `Engine(halo_depth=K+1)`) -- several slab contexts on ONE device, every
temporal-blocking group run without any exchange inside it, the D-row blocks copied between the contexts after each
group (what the library does over NCCL), against the oracle, bitwise.  (The layout and the plan are CPU-pinned:
tests/test_edge_emulation.py, tests/test_slabs.py.)"""
import os

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC
from fdtd_em_solver_b200 import slab as SL

pytestmark = [pytest.mark.gpu]           # first B200 run: round 2 (green); over NCCL: profiles/r2_slab_check_n2.log


def _eligible(setup):
    """stepK_edge does not take a TF/SF "right" line at column 0 (python's h[:, col-1] wraps)."""
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    return setup


def _copy_blocks(engines):
    import torch
    for r, e in enumerate(engines):
        D = e.halo_depth
        mine = e.generation_buffers()
        if r + 1 < len(engines):
            below = engines[r + 1]
            for a, b in zip(mine, below.generation_buffers()):
                b[below.row_begin - D - below.row_off:below.row_begin - below.row_off].copy_(
                    a[e.row_end - D - e.row_off:e.row_end - e.row_off])              # my last D rows -> its top halo
                a[e.row_end - e.row_off:e.row_end + D - e.row_off].copy_(
                    b[below.row_begin - below.row_off:below.row_begin + D - below.row_off])   # its first D rows -> my bottom halo
    torch.cuda.synchronize()


@pytest.mark.parametrize("world,K,shape,tile_rows", [(2, 4, (300, 700), 16), (3, 6, (400, 900), 16), (2, 2, (97, 300), 8),
                                                    (4, 6, (700, 2100), 48), (3, 8, (500, 1300), 24)])
def test_deep_halo_slabs_on_one_gpu_match_the_oracle(world, K, shape, tile_rows):
    from fdtd_em_solver_b200.engine import Engine
    rows, cols = shape
    n_steps = 3 * K + 2
    f, setup, times = SC.random_case(rows, cols, 11, walls=(False, True, False, False), n_steps=n_steps)
    setup = _eligible(setup)
    S = setup.courant
    engines = []
    for b, e_ in SL.split_rows(rows, world):
        e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), reflect=setup.reflect_walls, row_begin=b, row_end=e_,
                   temporal=K, tile_rows=tile_rows, halo_depth=K + 1)
        e.upload_coefficients(setup.eps_coef, setup.mu_coef)
        e.set_sources(setup.pulses, times)
        e.set_reflectors(setup.reflectors)
        e.set_fields(f.h, f.ex, f.ey)
        engines.append(e)
    n = 0
    while n < n_steps:
        g = min(K, n_steps - n) if n_steps - n >= K else 1        # whole groups, then single steps
        for e in engines:
            e.run(n, g)
        _copy_blocks(engines)
        n += g
    R.run(f, setup, times)
    h, ex, ey = np.zeros_like(f.h), np.zeros_like(f.ex), np.zeros_like(f.ey)
    for e in engines:
        e.get_fields(out={"h": h, "ex": ex, "ey": ey})
    assert np.array_equal(h, f.h) and np.array_equal(ex, f.ex) and np.array_equal(ey, f.ey)
    launches = [e.stats()["kernel_launches"] for e in engines]
    # per group: the plain kernel + one stepK_edge launch per tile class (<= 5) in the near and in the far half; one per
    # single step
    assert max(launches) <= 3 * (1 + 2 * 5) + 2 + 4
    for e in engines:
        e.close()
