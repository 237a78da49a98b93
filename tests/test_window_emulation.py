"""CPU check of frame_window's K-call window algorithm (fdtd_em_solver_b200/csrc/window.cuh) against the oracle.

The kernel advances the FRAME tiles of a temporal-blocking group (walls, pulse lines, reflectors, probes) K calls in
shared memory from one load of "tile + (K+1) cells of apron".  Its per-cell phase functions are __host__ __device__;
tests/host_emul/window_emul.cu runs them with plain loops, so the algorithm is pinned bit-for-bit here without a GPU:

  * K = 1..8 with every tile covering the grid (1x1, 2x3, 5x32, 8x64 tiles), all wall combinations, all pulse kinds,
    reflectors, probes: generation B must equal the oracle after K calls EVERYWHERE;
  * the apron K+1 is tight: K cells are not enough;
  * with the real frame plan of the library (fdtd_plan_frame): exactly the frame tiles are written, with the oracle's
    values, the rest of generation B is left to the K-step kernel.
"""
import ctypes as C
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

import resident_emul as E

pytestmark = pytest.mark.skipif(not E.available(), reason="nvcc not found (the emulation is compiled from window.cuh)")

WALLS = list(itertools.product([False, True], repeat=4))


def _eligible(setup):
    """A TF/SF "right" line at column 0 (python's h[:, col-1] wraps to the far column) is outside the window algorithm;
    the library keeps such scenes on the masked single-step passes."""
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]
    return setup


def _all_tiles(rows, cols, tr, tc):
    return [(tx, ty) for ty in range(-(-(rows + 1) // tr)) for tx in range(-(-(cols + 1) // tc))]


def _check(got, f, where, only=None):
    for name, a, b in zip(("h", "ex", "ey"), got[:3], (f.h, f.ex, f.ey)):
        if only is not None:
            m = only[name]
            a, b = a[m], b[m]
        if not np.array_equal(a, b):
            bad = np.argwhere(a != b)
            raise AssertionError("%s differs %s: %d cells, first %s" % (name, where, len(bad), bad[0]))


@pytest.mark.parametrize("K", [1, 2, 3, 4, 5, 6, 7, 8])
@pytest.mark.parametrize("tile", [(1, 1), (2, 3), (5, 32), (8, 64)])
def test_k_calls_every_tile_bitwise(K, tile):
    tr, tc = tile
    shapes = [(3, 3), (9, 5), (23, 41)] if tile in ((1, 1), (2, 3)) else [(9, 5), (40, 70), (67, 131)]
    for shape in shapes:
        rows, cols = shape
        for seed in (1, 2):
            walls = WALLS[(seed * 5 + rows + cols + K) % 16]
            f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=K + 3)
            setup = _eligible(setup)
            cells = [(0, 0), (rows - 1, cols - 1), (rows // 2, cols // 3)]
            f.raw_probes = [(i, j, []) for i, j in cells]
            first = seed                                      # the launch does not start at call 0
            for t in times[:first]:
                R.leapfrog(f, setup, t)
            for c in f.raw_probes:
                c[2].clear()
            got = E.window_emulate(f, setup, times, first, K, _all_tiles(rows, cols, tr, tc), tr, tc, order=(seed + K) % 3,
                                   probe_cells=cells)
            for t in times[first:first + K]:
                R.leapfrog(f, setup, t)
            _check(got, f, "K %d tile %s shape %s seed %d walls %s" % (K, tile, shape, seed, walls))
            for c, (i, j, series) in enumerate(f.raw_probes):
                assert np.array_equal(got[3][:, c, :], np.array(series))


@pytest.mark.parametrize("walls", WALLS)
def test_all_wall_combinations(walls):
    for shape, tile, K in (((6, 6), (1, 1), 4), ((17, 12), (2, 3), 6), ((12, 70), (8, 64), 6)):
        f, setup, times = SC.random_case(shape[0], shape[1], 11, walls=walls, n_steps=K)
        setup = _eligible(setup)
        got = E.window_emulate(f, setup, times, 0, K, _all_tiles(shape[0], shape[1], *tile), tile[0], tile[1])
        R.run(f, setup, times)
        _check(got, f, "walls %s shape %s" % (walls, shape))


def test_the_apron_is_tight():
    """K cells of apron instead of K+1 must go wrong somewhere (1x1 tiles at an absorbing wall: the first hop inward
    from a wall is two cells)."""
    K, wrong = 4, 0
    for seed in range(4):
        f, setup, times = SC.random_case(9, 9, seed, walls=(False, False, False, False), n_steps=K)
        setup = _eligible(setup)
        got = E.window_emulate(f, setup, times, 0, K, _all_tiles(9, 9, 1, 1), 1, 1, apron=K)
        ok = E.window_emulate(f, setup, times, 0, K, _all_tiles(9, 9, 1, 1), 1, 1, apron=K + 1)
        R.run(f, setup, times)
        _check(ok, f, "apron K+1")
        wrong += int(not all(np.array_equal(a, b) for a, b in zip(got[:3], (f.h, f.ex, f.ey))))
    assert wrong > 0


def test_the_shrinking_active_box_is_tight():
    """Call m runs on the window shrunk by m cells (window_margin).  Never shrinking gives the same bits; shrinking by
    two cells per call must go wrong."""
    K, wrong = 5, 0
    for seed in range(4):
        f, setup, times = SC.random_case(11, 13, seed, walls=(False, False, False, False), n_steps=K)
        setup = _eligible(setup)
        tiles = _all_tiles(11, 13, 2, 3)
        rule = E.window_emulate(f, setup, times, 0, K, tiles, 2, 3)
        never = E.window_emulate(f, setup, times, 0, K, tiles, 2, 3, shrink=0)
        greedy = E.window_emulate(f, setup, times, 0, K, tiles, 2, 3, shrink=2)
        R.run(f, setup, times)
        _check(rule, f, "kernel rule")
        _check(never, f, "no shrink")
        wrong += int(not all(np.array_equal(a, b) for a, b in zip(greedy[:3], (f.h, f.ex, f.ey))))
    assert wrong > 0


def _frame_plan(rows, cols, K, tile_rows, block_warps, boxes, row_begin=0, row_end=None):
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, row_begin, rows if row_end is None else row_end
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, block_warps
    plan = L.FramePlan()
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 4))
    rc = lib.fdtd_plan_frame(C.byref(cfg), K, len(b), b.ctypes.data if len(b) else None, C.byref(plan), None, None)
    assert rc == 0
    masks = np.zeros((K, plan.frame_tiles_y, plan.frame_tiles_x), dtype=np.uint8)
    skip = np.zeros((max(plan.k_tiles_y, 1), plan.k_tiles_x), dtype=np.uint8)
    rc = lib.fdtd_plan_frame(C.byref(cfg), K, len(b), b.ctypes.data if len(b) else None, C.byref(plan),
                             skip.ctypes.data, masks.ctypes.data)
    assert rc == 0
    return plan, masks


def _boxes(setup, rows, cols, probe_cells):
    """The inclusive boxes fdtd2d.cu build_frame_maps hands to the planner (all pulses counted as live)."""
    BIG = 1 << 40
    out = []
    for p in setup.pulses:
        if p.direction is None:
            out.append((p.row, p.row, p.col, p.col))
        elif p.direction in ("right", "left"):
            out.append((-BIG, BIG, p.col, p.col))
        elif p.direction == "up":
            out.append((p.row + 1, p.row + 1, -BIG, BIG))
        else:
            out.append((p.row, p.row, -BIG, BIG))
    for rect in setup.reflectors:
        for shape in ((rows, cols), (rows + 1, cols), (rows, cols + 1)):
            r0, r1, c0, c1 = E.clip_rect(rect, *shape)
            if r1 > r0 and c1 > c0:
                out.append((r0, r1 - 1, c0, c1 - 1))
    for i, j in probe_cells:
        out.append((i, i, j, j))
    return out


@pytest.mark.parametrize("K,tile_rows", [(2, 8), (4, 16), (6, 48), (8, 56)])
def test_frame_tiles_of_the_library_plan(K, tile_rows):
    """The frame of a real plan: frame_window writes exactly the frame tiles, with the oracle's values after K calls."""
    rows, cols = 300, 700
    f, setup, times = SC.random_case(rows, cols, 4, walls=(False, True, False, False), n_steps=K + 2)
    setup = _eligible(setup)
    cells = [(rows // 2, cols // 2), (5, cols // 3)]
    f.raw_probes = [(i, j, []) for i, j in cells]
    plan, masks = _frame_plan(rows, cols, K, tile_rows, 1, _boxes(setup, rows, cols, cells))
    tr, tc = plan.frame_tile_rows, plan.frame_tile_cols
    tiles = [(tx, ty) for ty, tx in np.argwhere(masks[0] != 0)]
    assert 0 < len(tiles) < masks[0].size                     # there is a frame, and there are K-step tiles
    R.leapfrog(f, setup, times[0])
    for c in f.raw_probes:
        c[2].clear()
    got = E.window_emulate(f, setup, times, 1, K, tiles, tr, tc, order=K % 3, probe_cells=cells)
    for t in times[1:1 + K]:
        R.leapfrog(f, setup, t)
    covered = {}
    for name, arr in zip(("h", "ex", "ey"), (f.h, f.ex, f.ey)):
        m = np.zeros(arr.shape, dtype=bool)
        for tx, ty in tiles:
            m[ty * tr:(ty + 1) * tr, tx * tc:(tx + 1) * tc] = True
        covered[name] = m
    _check(got, f, "frame tiles, K %d" % K, only=covered)
    for name, a in zip(("h", "ex", "ey"), got[:3]):
        assert np.isnan(a[~covered[name]]).all()              # nothing outside the frame tiles was written
    for c, (i, j, series) in enumerate(f.raw_probes):
        assert np.array_equal(got[3][:, c, :], np.array(series))


def _frame_plan_split(rows, cols, K, tile_rows, block_warps, boxes, row_begin, row_end):
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, row_begin, row_end
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, block_warps
    plan = L.FramePlan()
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 4))
    bp = b.ctypes.data if len(b) else None
    assert lib.fdtd_plan_frame(C.byref(cfg), K, len(b), bp, C.byref(plan), None, None) == 0
    full = np.zeros((K, plan.frame_tiles_y, plan.frame_tiles_x), dtype=np.uint8)
    assert lib.fdtd_plan_frame(C.byref(cfg), K, len(b), bp, C.byref(plan), None, full.ctypes.data) == 0
    chain = np.zeros_like(full)
    n = C.c_int32(0)
    xy = np.zeros((plan.frame_tiles_y * plan.frame_tiles_x, 2), dtype=np.int32)
    assert lib.fdtd_plan_frame_split(C.byref(cfg), K, len(b), bp, C.byref(plan), None, chain.ctypes.data,
                                     C.byref(n), xy.ctypes.data) == 0
    return plan, full, chain, [tuple(t) for t in xy[:n.value]]


@pytest.mark.parametrize("K,tile_rows", [(2, 8), (4, 16), (6, 48)])
@pytest.mark.parametrize("n_slabs", [1, 2, 3])
def test_y_slabs_window_tiles_and_interface_chain(K, tile_rows, n_slabs):
    """Frame modes 1/2 on y-slabs: frame_window takes the frame tiles whose window lies inside the rows the slab owns
    (the halo rows hold NaN here: they must not be read), the masked passes keep exactly the interface strips."""
    rows, cols = 330, 500
    f, setup, times = SC.random_case(rows, cols, 7, walls=(False, False, True, False), n_steps=K + 1)
    setup = _eligible(setup)
    boxes = _boxes(setup, rows, cols, [])
    g = R.Fields(f.h.copy(), f.ex.copy(), f.ey.copy())
    for t in times[:K]:
        R.leapfrog(g, setup, t)
    edges = [rows * s // n_slabs for s in range(n_slabs + 1)]
    for b, e in zip(edges[:-1], edges[1:]):
        plan, full, chain, tiles = _frame_plan_split(rows, cols, K, tile_rows, 1, boxes, b, e)
        tr, tc = plan.frame_tile_rows, plan.frame_tile_cols
        win = np.zeros_like(full[0])
        for tx, ty in tiles:
            win[ty, tx] = 1
        assert np.array_equal(win | chain[0], full[0]) and not (win & chain[0]).any()      # a partition of the frame
        A = K + 1
        for tx, ty in tiles:                                  # every window inside the owned rows, or beyond a real wall
            ta, tb = b + ty * tr, min(b + (ty + 1) * tr, e + (1 if e == rows else 0))
            assert b == 0 or ta - A >= b
            assert e == rows or tb + A <= e
        if n_slabs == 1:
            assert not chain.any() and len(tiles) == int(full[0].sum())
        else:
            assert chain[0].any() and len(tiles) > 0
            interface_rows = [r for r in (b, e - 1) if 0 < r < rows - 1]
            for r in interface_rows:                          # the rows next to a slab interface stay on the passes
                assert chain[0][(r - b) // tr].all()
        got = E.window_emulate(f, setup, times, 0, K, tiles, tr, tc, row_begin=b, row_end=e, poison_halo=True)
        covered = {}
        for name, arr in zip(("h", "ex", "ey"), (g.h, g.ex, g.ey)):
            m = np.zeros(arr.shape, dtype=bool)
            for tx, ty in tiles:
                m[b + ty * tr:min(b + (ty + 1) * tr, arr.shape[0]), tx * tc:(tx + 1) * tc] = True
            if name != "ex" or e < rows:
                m[e:] = False
            covered[name] = m
        _check(got, g, "slab [%d,%d) K %d" % (b, e, K), only=covered)
        for name, a in zip(("h", "ex", "ey"), got[:3]):
            assert np.isnan(a[~covered[name]]).all()


@pytest.mark.parametrize("M", [2, 3, 4, 8])
@pytest.mark.parametrize("name,n_calls", [("waveguide", 61), ("tutorial", 45), ("mixed", 38)])
def test_resident_window_intervals_of_m_calls(name, n_calls, M):
    """resident_window (resident mode with M calls per grid barrier): the script scenes advanced interval by interval,
    8 x 64 tiles with M+1 cells of apron, a shorter last interval on the same window size, probes every call."""
    sc = SC.spec(name)
    f, setup, times, info = SC.restate(sc)
    rows = cols = info["size"]
    cells = [(i, j) for i, j, _ in f.probes] or [(rows // 2, cols // 2)]
    f.probes = []
    f.raw_probes = [(i, j, []) for i, j in cells]
    g = R.Fields(f.h.copy(), f.ex.copy(), f.ey.copy())
    records = []
    tiles = _all_tiles(rows, cols, 8, 64)
    for n in range(0, n_calls, M):
        K = min(M, n_calls - n)
        h, ex, ey, rec = E.window_emulate(g, setup, times, n, K, tiles, 8, 64, apron=M + 1, order=(n // M) % 3,
                                          probe_cells=cells)
        g = R.Fields(h, ex, ey)
        records.append(rec)
    for t in times[:n_calls]:
        R.leapfrog(f, setup, t)
    _check((g.h, g.ex, g.ey), f, "%s, %d calls per barrier" % (name, M))
    rec = np.concatenate(records)
    for c, (i, j, series) in enumerate(f.raw_probes):
        assert np.array_equal(rec[:, c, :], np.array(series))


@pytest.mark.parametrize("K,tile_rows,variant", [(2, 8, 0), (4, 16, 1), (6, 16, 1), (6, 48, 0), (8, 24, 1)])
def test_whole_group_k_step_kernel_plus_frame_window_covers_the_grid(K, tile_rows, variant):
    """One temporal-blocking group as fdtd2d.cu group_window launches it: the K-step kernel (stepK_stream / stepK_lean,
    device code run by the SIMT emulator) on the plain tiles and frame_window on the frame tiles, both from generation A
    into the same generation B.  Together they must write EVERY cell, with the oracle's values after K calls."""
    import stream_emul as SE
    if not SE.available():
        pytest.skip("needs g++ and the CUDA headers")
    rows, cols = 260, 420
    f, setup, times = SC.random_case(rows, cols, 5, walls=(False, True, False, False), n_steps=K + 1)
    setup = _eligible(setup)
    R.leapfrog(f, setup, times[0])                             # generation A = the state after call 0
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, 0, rows
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, 2
    boxes = np.ascontiguousarray(np.asarray(_boxes(setup, rows, cols, []), dtype=np.int64).reshape(-1, 4))
    plan = L.FramePlan()
    n = C.c_int32(0)
    assert lib.fdtd_plan_frame_split(C.byref(cfg), K, len(boxes), boxes.ctypes.data, C.byref(plan), None, None,
                                     C.byref(n), None) == 0
    skip = np.zeros((max(plan.k_tiles_y, 1), plan.k_tiles_x), dtype=np.uint8)
    xy = np.zeros((plan.frame_tiles_y * plan.frame_tiles_x, 2), dtype=np.int32)
    masks = np.zeros((K, plan.frame_tiles_y, plan.frame_tiles_x), dtype=np.uint8)
    assert lib.fdtd_plan_frame_split(C.byref(cfg), K, len(boxes), boxes.ctypes.data, C.byref(plan), skip.ctypes.data,
                                     masks.ctypes.data, C.byref(n), xy.ctypes.data) == 0
    assert not masks.any() and n.value > 0 and plan.k_tiles_y > 0 and not skip.all()
    tiles = [tuple(t) for t in xy[:n.value]]
    by_k = SE.stepK_stream(f, setup, times, 1, K, tile_rows, 2, plan.k_row_lo, plan.k_row_hi, skip, variant=variant)
    by_w = E.window_emulate(f, setup, times, 1, K, tiles, plan.frame_tile_rows, plan.frame_tile_cols)[:3]
    for t in times[1:1 + K]:
        R.leapfrog(f, setup, t)
    for name, a, b, want in zip(("h", "ex", "ey"), by_k, by_w, (f.h, f.ex, f.ey)):
        both = ~np.isnan(a) & ~np.isnan(b)
        assert np.array_equal(a[both], b[both]), name         # where they overlap they agree
        merged = np.where(np.isnan(a), b, a)
        assert not np.isnan(merged).any(), "%s: %d cells written by neither kernel" % (name, int(np.isnan(merged).sum()))
        assert np.array_equal(merged, want), name


def _frame_plan_halo(rows, cols, K, depth, tile_rows, block_warps, boxes, row_begin, row_end):
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, row_begin, row_end
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, block_warps
    plan = L.FramePlan()
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 4))
    bp = b.ctypes.data if len(b) else None
    n = C.c_int32(0)
    assert lib.fdtd_plan_frame_halo(C.byref(cfg), K, depth, len(b), bp, C.byref(plan), None, None, C.byref(n), None) == 0
    skip = np.zeros((max(plan.k_tiles_y, 1), plan.k_tiles_x), dtype=np.uint8)
    masks = np.zeros((K, plan.frame_tiles_y, plan.frame_tiles_x), dtype=np.uint8)
    xy = np.zeros((plan.frame_tiles_y * plan.frame_tiles_x, 2), dtype=np.int32)
    assert lib.fdtd_plan_frame_halo(C.byref(cfg), K, depth, len(b), bp, C.byref(plan), skip.ctypes.data, masks.ctypes.data,
                                    C.byref(n), xy.ctypes.data) == 0
    return plan, skip[:plan.k_tiles_y], masks, [tuple(t) for t in xy[:n.value]]


@pytest.mark.parametrize("K,tile_rows,variant", [(2, 8, 0), (4, 16, 1), (6, 16, 0), (6, 48, 1), (8, 24, 0)])
@pytest.mark.parametrize("n_slabs", [2, 3])
def test_deep_halo_slabs_need_no_exchange_inside_a_group(K, tile_rows, variant, n_slabs):
    """Deep halos (fdtd_set_halo_depth(K+1)): every slab stores K+1 valid rows of each neighbour and NOTHING else of
    them (NaN here).  The library's plan for that layout has no pass chain left -- the K-step tiles reach the slab's
    first and last owned row, frame_window takes every frame tile -- and the two kernels (device code of the K-step
    kernel through the SIMT emulator, the window algorithm through its emulation) write every owned cell with the
    oracle's values after K calls, reading no row beyond the halos."""
    import stream_emul as SE
    if not SE.available():
        pytest.skip("needs g++ and the CUDA headers")
    rows, cols, D = 264, 420, K + 1
    f, setup, times = SC.random_case(rows, cols, 7, walls=(False, True, False, False), n_steps=K + 1)
    setup = _eligible(setup)
    R.leapfrog(f, setup, times[0])                             # generation A = the state after call 0
    g = R.Fields(f.h.copy(), f.ex.copy(), f.ey.copy(), step=f.step)
    for t in times[1:1 + K]:
        R.leapfrog(g, setup, t)
    edges = [(rows * s // n_slabs) // 8 * 8 for s in range(n_slabs)] + [rows]
    boxes = _boxes(setup, rows, cols, [])
    for b, e in zip(edges[:-1], edges[1:]):
        plan, skip, masks, tiles = _frame_plan_halo(rows, cols, K, D, tile_rows, 2, boxes, b, e)
        assert not masks.any() and tiles                      # no interface strips on a pass chain: all frame tiles are window tiles
        if b > 0:
            assert plan.k_row_lo == b                         # K-step tiles start at the first owned row of an interface
        if e < rows:
            assert plan.k_row_hi > e - plan.frame_tile_rows   # ... and end at the last (up to frame-tile alignment)
        stored_lo, stored_hi = (b - D if b > 0 else 0), (e + D if e < rows else rows + 1)
        poisoned = R.Fields(f.h.copy(), f.ex.copy(), f.ey.copy(), step=f.step)
        for a in (poisoned.h, poisoned.ex, poisoned.ey):
            a[:stored_lo] = np.nan
            a[stored_hi:] = np.nan
        by_k = SE.stepK_stream(poisoned, setup, times, 1, K, tile_rows, 2, plan.k_row_lo, plan.k_row_hi, skip, variant=variant)
        by_w = E.window_emulate(f, setup, times, 1, K, tiles, plan.frame_tile_rows, plan.frame_tile_cols, row_begin=b,
                                row_end=e, halo_depth=D)[:3]
        for name, a, w, want in zip(("h", "ex", "ey"), by_k, by_w, (g.h, g.ex, g.ey)):
            hi = e + (1 if (name == "ex" and e == rows) else 0)
            a, w, want = a[b:hi], w[b:hi], want[b:hi]
            merged = np.where(np.isnan(a), w, a)
            assert not np.isnan(merged).any(), "%s slab [%d,%d): %d owned cells unwritten or poisoned" % (
                name, b, e, int(np.isnan(merged).sum()))
            assert np.array_equal(merged, want), (name, b, e)
