"""stepK_edge (fdtd_em_solver_b200/csrc/edge.cuh) -- the K-step register pipeline with walls, pulses of every kind,
reflectors and probes folded in as tile predicates -- executed on the CPU by the SIMT emulator (tests/host_emul/simt.h,
the DEVICE code itself, one warp at a time) and compared bitwise with the oracle.

Here EVERY tile of the index space is an edge tile, so walls on all sides, corners, sources and reflectors anywhere, and
tiles that hold both side walls at once all go through the predicates; the arrays are padded with NaN outside the
fields, so anything read from a non-existent cell that reached a stored value would show."""
import itertools

import numpy as np
import pytest

from oracle import restated as R, scenarios as SC

import stream_emul as E

pytestmark = pytest.mark.skipif(not E.available(), reason="needs g++ and the CUDA headers")

WALLS = list(itertools.product([False, True], repeat=4))


def geometry(K, V=2):
    margin = (K + V - 1) // V
    return margin, (32 - 2 * margin) * V


def all_tiles(rows, cols, K, heights):
    """(tx, ta, tb) covering rows [0, rows] x cols [0, cols] of the index space; tile heights cycle through `heights`."""
    margin, out = geometry(K)
    ntx = (cols + 1 + out - 1) // out
    while ntx > 1 and (ntx - 2) * out - margin * 2 + 64 > cols:      # the tile before already loads (and so owns) up to column NC
        ntx -= 1
    tiles, ta, k = [], 0, 0
    while ta < rows + 1:
        tb = min(ta + heights[k % len(heights)], rows + 1)
        tiles += [(tx, ta, tb) for tx in range(ntx)]
        ta, k = tb, k + 1
    return tiles


def drop_wrapping_line(setup):
    """A TF/SF "right" line at column 0 wraps to the far column (App. A-Q4): such scenes never reach the K-step path."""
    setup.pulses = [p for p in setup.pulses if not (p.direction == "right" and p.col == 0)]


def check(got, f, where):
    for name, a, b in zip(("h", "ex", "ey"), got, (f.h, f.ex, f.ey)):
        if not np.array_equal(a, b):
            bad = np.argwhere(~((a == b) | (np.isnan(a) & np.isnan(b))))
            raise AssertionError("%s differs %s: %d cells, first %s (got %r want %r)" % (
                name, where, len(bad), bad[0], a[tuple(bad[0])], b[tuple(bad[0])]))


@pytest.mark.parametrize("K", [1, 2, 3, 4, 6, 8])
@pytest.mark.parametrize("shape,heights", [((9, 5), (3, 2)), ((33, 40), (7, 5, 12)), ((40, 130), (9,)), ((70, 61), (70,)),
                                           ((26, 200), (4, 11))])
def test_edge_tiles_cover_whole_grids_bitwise(shape, heights, K):
    rows, cols = shape
    for seed in (0, 1, 2, 5):
        walls = WALLS[(seed * 7 + rows + cols + K) % 16]
        f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=2 * K + 1)
        drop_wrapping_line(setup)
        tiles = all_tiles(rows, cols, K, heights)
        for n in (0, K):                                   # two consecutive groups: different pulse windows
            got, _ = E.stepK_edge(f, setup, times, n, K, tiles)
            for t in times[n:n + K]:
                R.leapfrog(f, setup, t)
            check(got, f, "shape %s seed %d walls %s K %d group at %d" % (shape, seed, walls, K, n))


@pytest.mark.parametrize("walls", WALLS)
def test_every_wall_combination(walls):
    f, setup, times = SC.random_case(21, 75, 3, walls=walls, n_steps=6)
    drop_wrapping_line(setup)
    got, _ = E.stepK_edge(f, setup, times, 0, 6, all_tiles(21, 75, 6, (8, 5)))
    R.run(f, setup, times)
    check(got, f, "walls %s" % (walls,))


def test_probes_record_every_call_of_the_group():
    rows, cols, K = 30, 90, 5
    f, setup, times = SC.random_case(rows, cols, 4, walls=(False, True, False, False), n_steps=K)
    drop_wrapping_line(setup)
    cells = [(0, 0), (0, 17), (rows - 1, cols - 1), (12, 51), (29, 0), (7, 52), (1, 1)]
    got, rec = E.stepK_edge(f, setup, times, 0, K, all_tiles(rows, cols, K, (6, 9)), probes=cells)
    want = np.zeros((K, len(cells), 3))
    for m, t in enumerate(times):
        R.leapfrog(f, setup, t)
        for k, (i, j) in enumerate(cells):
            want[m, k] = (f.h[i, j], f.ex[i, j], f.ey[i, j])
    check(got, f, "probe scene")
    assert np.array_equal(rec, want)


def test_only_listed_tiles_are_written():
    rows, cols, K = 40, 130, 4
    f, setup, times = SC.random_case(rows, cols, 6, n_steps=K)
    drop_wrapping_line(setup)
    _, out = geometry(K)
    tiles = [(1, 10, 19), (0, 0, 6), (2, 33, 41)]
    got, _ = E.stepK_edge(f, setup, times, 0, K, tiles)
    R.run(f, setup, times)
    for name, a, b in zip(("h", "ex", "ey"), got, (f.h, f.ex, f.ey)):
        m = np.zeros(a.shape, dtype=bool)
        for tx, ta, tb in tiles:
            m[ta:tb, tx * out:(tx + 1) * out] = True            # (none of these tiles reaches the right wall)
        assert np.array_equal(a[m], b[m]), name
        assert np.isnan(a[~m]).all(), name


# ---- the library's own tile plan (fdtd_plan_edge): plain tiles through stepK_lean / stepK_stream, all the others through
# ---- stepK_edge, together one whole group ------------------------------------------------------------------------------
def _plan_edge(rows, cols, K, tile_rows, boxes=(), row_begin=0, row_end=None, halo_depth=1):
    """-> plan, skip map, tiles (tx, ta, tb, next-to-an-interface flag, class).  A box is (r0, r1, c0, c1[, is_pulse])."""
    import ctypes as C
    from fdtd_em_solver_b200 import _lib as L
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, row_begin, rows if row_end is None else row_end
    cfg.dtype, cfg.tile_rows, cfg.block_warps = L.F64, tile_rows, 1
    plan = L.FramePlan()
    b = np.ascontiguousarray(np.asarray([q[:4] for q in boxes], dtype=np.int64).reshape(-1, 4))
    kinds = np.ascontiguousarray([q[4] for q in boxes if len(q) > 4], dtype=np.int32)
    kp = kinds.ctypes.data if len(boxes) and len(kinds) == len(boxes) else None
    n = C.c_int32(0)
    assert lib.fdtd_plan_edge(C.byref(cfg), K, halo_depth, len(b), b.ctypes.data, kp, C.byref(plan), None, C.byref(n), None, 0) == 0
    skip = np.ones((max(plan.k_tiles_y, 1), plan.k_tiles_x), dtype=np.uint8)
    tiles = np.zeros((max(n.value, 1), 4), dtype=np.int32)
    assert lib.fdtd_plan_edge(C.byref(cfg), K, halo_depth, len(b), b.ctypes.data, kp, C.byref(plan), skip.ctypes.data, C.byref(n),
                              tiles.ctypes.data, len(tiles)) == 0
    tiles = tiles[:n.value]
    return plan, skip, np.column_stack([tiles[:, :3], tiles[:, 3] & 1, tiles[:, 3] >> 8])


def _boxes(setup, rows, cols):
    """The modifier boxes build_frame_maps hands to the planner (all pulses live)."""
    from resident_emul import clip_rect
    big = 1 << 40
    out = []
    for p in setup.pulses:
        if p.direction is None: out.append((p.row, p.row, p.col, p.col, 1))
        elif p.direction in ("right", "left"): out.append((-big, big, p.col, p.col, 1))
        elif p.direction == "up": out.append((p.row + 1, p.row + 1, -big, big, 1))
        else: out.append((p.row, p.row, -big, big, 1))
    for rect in setup.reflectors:
        for shape in ((rows, cols), (rows + 1, cols), (rows, cols + 1)):
            r0, r1, c0, c1 = clip_rect(rect, *shape)
            if r1 > r0 and c1 > c0:
                out.append((r0, r1 - 1, c0, c1 - 1, 0))
    return out


@pytest.mark.parametrize("K,tile_rows", [(2, 8), (3, 12), (4, 16), (6, 16), (6, 48), (7, 9), (8, 24)])
@pytest.mark.parametrize("shape", [(100, 230), (61, 700)])
def test_plain_plus_edge_tiles_are_one_whole_group(shape, K, tile_rows):
    rows, cols = shape
    seen_classes = set()
    for seed, walls in ((9, (False,) * 4), (2, (True, False, False, True)), (7, (False, True, True, False))):
        f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=K)
        drop_wrapping_line(setup)
        plan, skip, tiles = _plan_edge(rows, cols, K, tile_rows, _boxes(setup, rows, cols))
        # every cell of the index space belongs to exactly one tile
        margin, out = geometry(K)
        owner = np.zeros((rows + 1, cols + 1), dtype=np.int32)
        for ty, tx in np.argwhere(skip[:plan.k_tiles_y] == 0):
            ta = plan.k_row_lo + ty * plan.k_tile_rows
            owner[ta:min(ta + plan.k_tile_rows, plan.k_row_hi), tx * out:(tx + 1) * out] += 1
        for tx, ta, tb, _, _ in tiles:
            reach = tx * out - margin * 2 + 64 > cols
            owner[ta:tb, tx * out:(cols + 1 if reach else (tx + 1) * out)] += 1
        assert (owner == 1).all(), np.argwhere(owner != 1)[:5]
        seen_classes.update(int(t[4]) for t in tiles)
        # every tile through the instantiation of ITS class: a feature the planner left out must not have been needed
        got_e, _ = E.stepK_edge(f, setup, times, 0, K, [(t[0], t[1], t[2], t[4]) for t in tiles])
        got_p = None
        if plan.k_tiles_y > 0 and not skip.all():
            quiet = R.Setup(eps_coef=setup.eps_coef, mu_coef=setup.mu_coef, courant=setup.courant,
                            reflect_walls=setup.reflect_walls, pulses=[], reflectors=[])
            got_p = E.stepK_stream(f, quiet, times, 0, K, tile_rows, 1, plan.k_row_lo, plan.k_row_hi, skip, variant=1)
        R.run(f, setup, times)
        for k, (name, want) in enumerate(zip(("h", "ex", "ey"), (f.h, f.ex, f.ey))):
            a = got_e[k].copy()
            if got_p is not None:
                both = ~np.isnan(a) & ~np.isnan(got_p[k])
                assert not both.any(), name                      # the two launches write disjoint cells
                a = np.where(np.isnan(a), got_p[k], a)
            assert np.array_equal(a, want), (name, seed, np.argwhere(a != want)[:4])
    assert seen_classes <= {1, 2, 4, 6, 15}, seen_classes


@pytest.mark.parametrize("K,tile_rows", [(2, 8), (4, 16), (6, 16), (8, 24)])
@pytest.mark.parametrize("line", [None, "right", "left", "down"])
def test_every_tile_class_is_planned_and_exact(K, tile_rows, line):
    """A scene shaped like the C5 recipe (walls, one point source, at most one TF/SF line, one reflector, one probe): the
    planner gives the wall columns, the wall bands, the source tiles and the TF/SF column next to a wall their own cheap
    instantiations, and the group is still the oracle's, bit for bit."""
    rows, cols = 150, 520
    rng = np.random.default_rng(K)
    f, setup, times = SC.random_case(rows, cols, 11, walls=(False,) * 4, n_steps=K, with_sources=False, with_rects=False)
    setup.pulses = [R.PulseSpec("table", None, 70, 300, -1.0, 1.0, rng.standard_normal(K))]
    if line:
        setup.pulses.append(R.PulseSpec("table", line, 40, 7, -1.0, 1.0, rng.standard_normal(K)))
    setup.reflectors = [[100, 110, 400, 430]]
    probes = [(30, 200)]
    boxes = _boxes(setup, rows, cols) + [(30, 30, 200, 200, 0)]
    plan, skip, tiles = _plan_edge(rows, cols, K, tile_rows, boxes)
    classes = {int(t[4]) for t in tiles}
    assert classes == ({1, 2, 4, 6, 15} if line in ("right", "left") else {1, 2, 4, 15} if line is None else {1, 2, 4, 6, 15}), classes
    got_e, rec = E.stepK_edge(f, setup, times, 0, K, [(t[0], t[1], t[2], t[4]) for t in tiles], probes=probes)
    quiet = R.Setup(eps_coef=setup.eps_coef, mu_coef=setup.mu_coef, courant=setup.courant, reflect_walls=setup.reflect_walls,
                    pulses=[], reflectors=[])
    got_p = E.stepK_stream(f, quiet, times, 0, K, tile_rows, 1, plan.k_row_lo, plan.k_row_hi, skip, variant=1)
    want_rec = np.zeros((K, 1, 3))
    for m, t in enumerate(times):
        R.leapfrog(f, setup, t)
        want_rec[m, 0] = (f.h[30, 200], f.ex[30, 200], f.ey[30, 200])
    for k, (name, want) in enumerate(zip(("h", "ex", "ey"), (f.h, f.ex, f.ey))):
        a = np.where(np.isnan(got_e[k]), got_p[k], got_e[k])
        assert np.array_equal(a, want), (name, np.argwhere(a != want)[:4])
    assert np.array_equal(rec, want_rec)


@pytest.mark.parametrize("K", [2, 4, 6, 8])
def test_edge_plan_on_slabs_with_deep_halos(K):
    """Per slab: the tiles cover exactly the owned rows; no tile marches through a row the slab does not store."""
    rows, cols, D = 300, 200, K + 1
    margin, out = geometry(K)
    for world in (2, 3):
        from fdtd_em_solver_b200.slab import split_rows
        for b, e in split_rows(rows, world):
            plan, skip, tiles = _plan_edge(rows, cols, K, 16, [(40, 40, 70, 70)], row_begin=b, row_end=e, halo_depth=D)
            lo, hi = b, e + (1 if e == rows else 0)
            owner = np.zeros((rows + 1, cols + 1), dtype=np.int32)
            first, last = max(b - D, 0) if b > 0 else 0, (e + D) if e < rows else rows
            for ty, tx in np.argwhere(skip[:plan.k_tiles_y] == 0):
                ta = plan.k_row_lo + ty * plan.k_tile_rows
                tb = min(ta + plan.k_tile_rows, plan.k_row_hi)
                owner[ta:tb, tx * out:(tx + 1) * out] += 1
                assert ta - K >= first and tb + K - 1 <= last
            for tx, ta, tb, near, _ in tiles:
                reach = tx * out - margin * 2 + 64 > cols
                owner[ta:tb, tx * out:(cols + 1 if reach else (tx + 1) * out)] += 1
                assert max(ta - K - 1, 0) >= first and min(tb + K, rows) <= last
                assert bool(near) == ((b > 0 and ta < b + D) or (e < rows and tb > e - D))
            assert (owner[lo:hi] == 1).all() and (owner[:lo] == 0).all() and (owner[hi:] == 0).all()


# ---- split field storage (dtype FDTD_F32X): the same kernel with SPLIT = 1, held to oracle/split_storage.py ---------------
@pytest.mark.parametrize("K", [1, 2, 3, 5, 8])
@pytest.mark.parametrize("shape,heights", [((33, 40), (7, 5, 12)), ((40, 130), (9,)), ((26, 200), (4, 11))])
def test_split_storage_group_of_k_calls_has_the_bits_of_k_stored_calls(shape, heights, K):
    """fp32 hi + bf16 lo field planes, float64 coefficients and arithmetic: a launch that advances K calls must leave what
    K calls with a store and a load between them leave -- every value the next call reads passes through split_round."""
    from oracle import split_storage as SS
    rows, cols = shape
    for seed in (0, 1, 5):
        walls = WALLS[(seed * 7 + rows + cols + K) % 16]
        f, setup, times = SC.random_case(rows, cols, seed, walls=walls, n_steps=2 * K + 1)
        drop_wrapping_line(setup)
        SS.store(f)                                            # the device holds the fields in split storage
        cells = [(0, 0), (rows // 2, cols // 3), (rows - 1, cols - 1)]
        tiles = all_tiles(rows, cols, K, heights)
        for n in (0, K):
            got, rec = E.stepK_edge(f, setup, times, n, K, tiles, probes=cells, split=True)
            want_rec = np.zeros((K, len(cells), 3))
            for m, t in enumerate(times[n:n + K]):
                SS.leapfrog(f, setup, t)
                for k, (i, j) in enumerate(cells):
                    want_rec[m, k] = (f.h[i, j], f.ex[i, j], f.ey[i, j])
            check(got, f, "split storage: shape %s seed %d walls %s K %d group at %d" % (shape, seed, walls, K, n))
            assert np.array_equal(rec, want_rec)


@pytest.mark.parametrize("K,tile_rows", [(2, 8), (3, 12), (5, 16), (8, 24), (8, 9)])
def test_split_storage_group_on_the_library_plan_plain_and_edge_tiles(K, tile_rows):
    """What fdtd_run launches for a group under split storage: the plain tiles of the frame plan through stepK_lean_split
    (kernels.cuh), every other tile through stepK_edge<.., split>; together they write every cell once, with the bits of K
    stored calls -- two consecutive groups, so the second one starts from stored values the first one produced."""
    from oracle import split_storage as SS
    rows, cols = 150, 520
    rng = np.random.default_rng(K)
    f, setup, times = SC.random_case(rows, cols, 11, walls=(False, True, False, False), n_steps=2 * K, with_sources=False,
                                     with_rects=False)
    setup.pulses = [R.PulseSpec("table", None, 70, 300, -1.0, 1.0, rng.standard_normal(2 * K)),
                    R.PulseSpec("table", "right", 40, 7, -1.0, 1.0, rng.standard_normal(2 * K))]
    setup.reflectors = [[100, 110, 400, 430]]
    probes = [(30, 200)]
    boxes = _boxes(setup, rows, cols) + [(30, 30, 200, 200, 0)]
    plan, skip, tiles = _plan_edge(rows, cols, K, tile_rows, boxes)
    assert (skip == 0).any() and len(tiles) > 0                 # the group really has both kinds of tiles
    quiet = R.Setup(eps_coef=setup.eps_coef, mu_coef=setup.mu_coef, courant=setup.courant, reflect_walls=setup.reflect_walls,
                    pulses=[], reflectors=[])
    SS.store(f)
    for n in (0, K):
        got_e, rec = E.stepK_edge(f, setup, times, n, K, [(t[0], t[1], t[2]) for t in tiles], probes=probes, split=True)
        got_p = E.stepK_lean_split(f, quiet, times, n, K, tile_rows, 1, plan.k_row_lo, plan.k_row_hi, skip)
        want_rec = np.zeros((K, 1, 3))
        for m, t in enumerate(times[n:n + K]):
            SS.leapfrog(f, setup, t)
            want_rec[m, 0] = (f.h[30, 200], f.ex[30, 200], f.ey[30, 200])
        for k, (name, want) in enumerate(zip(("h", "ex", "ey"), (f.h, f.ex, f.ey))):
            both = ~np.isnan(got_e[k]) & ~np.isnan(got_p[k])
            assert not both.any(), (name, "written twice", np.argwhere(both)[:4])
            a = np.where(np.isnan(got_e[k]), got_p[k], got_e[k])
            assert np.array_equal(a, want), (name, n, np.argwhere(a != want)[:4])
        assert np.array_equal(rec, want_rec)


@pytest.mark.parametrize("variant", [2, 3], ids=["stepK_lean_umu", "stepK_skew_umu"])
@pytest.mark.parametrize("K,tile_rows", [(3, 12), (5, 16), (8, 24), (8, 7)])
def test_whole_group_with_the_uniform_mu_kernels(K, tile_rows, variant):
    """The group fdtd_run launches when the mu_coef plane is one value: the plain tiles through stepK_lean_umu or
    stepK_skew_umu (value from the kernel parameters; the skewed pipeline runs extra drain iterations per tile), every other
    tile through stepK_edge, which still reads the plane.  Two consecutive groups, every cell written once, oracle's bits."""
    rows, cols = 120, 410
    f, setup, times = SC.random_case(rows, cols, 13, walls=(False, True, False, False), n_steps=2 * K)
    drop_wrapping_line(setup)
    setup.mu_coef = np.full_like(setup.mu_coef, setup.mu_coef[1, 1])
    plan, skip, tiles = _plan_edge(rows, cols, K, tile_rows, _boxes(setup, rows, cols))
    assert plan.k_tiles_y > 0 and not skip.all() and len(tiles) > 0
    quiet = R.Setup(eps_coef=setup.eps_coef, mu_coef=setup.mu_coef, courant=setup.courant, reflect_walls=setup.reflect_walls,
                    pulses=[], reflectors=[])
    for n in (0, K):
        got_e, _ = E.stepK_edge(f, setup, times, n, K, [(t[0], t[1], t[2], t[4]) for t in tiles])
        got_p = E.stepK_stream(f, quiet, times, n, K, tile_rows, 1, plan.k_row_lo, plan.k_row_hi, skip, variant=variant)
        for t in times[n:n + K]:
            R.leapfrog(f, setup, t)
        for k, (name, want) in enumerate(zip(("h", "ex", "ey"), (f.h, f.ex, f.ey))):
            both = ~np.isnan(got_e[k]) & ~np.isnan(got_p[k])
            assert not both.any(), name
            a = np.where(np.isnan(got_e[k]), got_p[k], got_e[k])
            assert np.array_equal(a, want), (name, n, np.argwhere(a != want)[:4])
