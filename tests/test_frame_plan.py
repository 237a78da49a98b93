"""CPU: the temporal-blocking frame plan (pure host logic of libfdtd2d, fdtd_plan_frame) -- invariants that make
K-step tiles + masked frame passes equal to K single steps, checked without a GPU on many geometries."""
import ctypes as C
import itertools

import numpy as np
import pytest

from fdtd_em_solver_b200 import _lib as L


def plan(rows, cols, K, tile_rows, boxes=(), row_begin=0, row_end=None, dtype=L.F64, warps=1):
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end = rows, cols, row_begin, rows if row_end is None else row_end
    cfg.dtype, cfg.tile_rows, cfg.block_warps = dtype, tile_rows, warps
    b = np.ascontiguousarray(np.asarray(boxes, dtype=np.int64).reshape(-1, 4))
    p = L.FramePlan()
    assert lib.fdtd_plan_frame(C.byref(cfg), K, len(b), C.c_void_p(b.ctypes.data), C.byref(p), None, None) == 0
    skip = np.zeros((max(p.k_tiles_y, 1), p.k_tiles_x), dtype=np.uint8)
    masks = np.zeros((K, p.frame_tiles_y, p.frame_tiles_x), dtype=np.uint8)
    assert lib.fdtd_plan_frame(C.byref(cfg), K, len(b), C.c_void_p(b.ctypes.data), C.byref(p),
                               C.c_void_p(skip.ctypes.data), C.c_void_p(masks.ctypes.data)) == 0
    return cfg, p, skip[:p.k_tiles_y], masks


def check(rows, cols, K, tile_rows, boxes=(), row_begin=0, row_end=None, dtype=L.F64):
    cfg, p, skip, masks = plan(rows, cols, K, tile_rows, boxes, row_begin, row_end, dtype)
    row_end = cfg.row_end
    lo, hi = row_begin, row_end + (1 if row_end == rows else 0)          # index rows this slab writes
    ncol = cols + 1
    # who writes each cell of the (hi-lo) x (cols+1) index space?
    by_k = np.zeros((hi - lo, ncol), dtype=bool)
    for ty in range(p.k_tiles_y):
        ta = p.k_row_lo + ty * p.k_tile_rows
        tb = min(ta + p.k_tile_rows, p.k_row_hi)
        for tx in range(p.k_tiles_x):
            if skip[ty, tx]:
                continue
            c0 = tx * p.k_tile_cols
            jw = c0 - p.k_margin_cols
            # (2) everything a K-step tile touches is an owned, plain interior cell and far from every box
            assert ta - K >= max(row_begin, 0) and ta - (K - 1) >= 1 and tb + K - 2 <= rows - 2 and tb + K - 1 <= row_end - 1
            assert jw >= 1 and jw + p.k_load_cols - 1 <= cols - 2
            for r0, r1, b0, b1 in boxes:
                assert not (r0 <= tb + K - 1 and r1 >= ta - K and b0 <= jw + p.k_load_cols - 1 and b1 >= jw), (ty, tx)
            assert not by_k[ta - lo:tb - lo, c0:c0 + p.k_tile_cols].any()      # K tiles never overlap
            by_k[ta - lo:tb - lo, c0:min(c0 + p.k_tile_cols, ncol)] = True
    frame = np.kron(masks[0], np.ones((p.frame_tile_rows, p.frame_tile_cols), dtype=np.uint8))[:hi - lo, :ncol] > 0
    # (1) every cell is written by the K-step kernel or by the last frame pass (both is allowed: same values)
    assert (by_k | frame).all()
    # the walls and every box are frame
    assert frame[0 if row_begin == 0 else slice(0, 0)].all() and frame[:, 0].all() and frame[:, cols - 1:].all()
    for r0, r1, b0, b1 in boxes:
        rr0, rr1 = max(r0, lo), min(r1, hi - 1)
        if rr0 <= rr1:
            assert frame[rr0 - lo:rr1 - lo + 1, max(b0, 0):min(b1, cols) + 1].all()
    # (3) ring d contains ring d-1 dilated by one tile; the ring used by pass p is thick enough for its K-p cells
    for d in range(1, K):
        m = masks[d - 1].astype(bool)
        grown = m.copy()
        for dy, dx in itertools.product((-1, 0, 1), repeat=2):
            shifted = np.zeros_like(m)
            ys = slice(max(dy, 0), m.shape[0] + min(dy, 0)); yd = slice(max(-dy, 0), m.shape[0] + min(-dy, 0))
            xs = slice(max(dx, 0), m.shape[1] + min(dx, 0)); xd = slice(max(-dx, 0), m.shape[1] + min(-dx, 0))
            shifted[yd, xd] = m[ys, xs]
            grown |= shifted
        assert np.array_equal(masks[d].astype(bool), grown)
    for pth in range(1, K):
        rings = -(-(K - pth) // p.frame_tile_rows)
        assert rings * min(p.frame_tile_rows, p.frame_tile_cols) >= K - pth and rings < K
    assert p.k_tile_rows % p.frame_tile_rows == 0 and (p.k_row_lo - lo) % p.frame_tile_rows == 0
    return p, skip, masks


@pytest.mark.parametrize("K", [2, 3, 4, 6, 8])
@pytest.mark.parametrize("rows,cols,tile_rows", [(200, 900, 8), (131, 517, 4), (64, 300, 16), (512, 3000, 48), (40, 130, 12)])
def test_plan_invariants_on_whole_grids(K, rows, cols, tile_rows):
    boxes = [(rows // 2, rows // 2, cols // 3, cols // 3), (-2 ** 40, 2 ** 40, 7, 7), (rows // 4, rows // 4, -2 ** 40, 2 ** 40),
             (10, 30, cols // 2, cols // 2 + 40)]
    check(rows, cols, K, tile_rows, boxes)
    check(rows, cols, K, tile_rows, [])
    check(rows, cols, K, tile_rows, boxes, dtype=L.F32)


@pytest.mark.parametrize("K", [2, 4, 6])
def test_plan_invariants_on_slabs(K):
    rows, cols = 700, 2000
    for b, e in [(0, 180), (180, 350), (350, 530), (530, 700)]:
        p, skip, masks = check(rows, cols, K, 16, [(400, 400, 100, 100), (-2 ** 40, 2 ** 40, 7, 7)], row_begin=b, row_end=e)
        # the interface rows of a slab are frame: K-step tiles keep K rows away from both ends of the slab
        assert p.k_row_lo >= b + K and p.k_row_hi <= e - K


def test_c5_slab_plan_numbers():
    """The bench workload: only thin bands are frame (numbers quoted in DESIGN.md)."""
    _, p, skip, masks = plan(8192, 65536, 6, 48, [(2, 2, 21845, 21845), (-2 ** 40, 2 ** 40, 7, 7)])
    assert (p.k_tile_cols, p.k_load_cols, p.k_margin_cols, p.frame_tile_rows, p.frame_tile_cols) == (52, 64, 6, 8, 64)
    assert p.k_row_lo == 8 and p.k_row_hi == 8184
    frac_frame = masks[0].mean()
    assert 0.004 < frac_frame < 0.008 and int(masks[0].sum()) == 6159
    assert skip.mean() < 0.004


def test_deep_halo_layout_and_argument_checks():
    lib = L.load()
    cfg = L.Config()
    cfg.rows, cfg.cols, cfg.row_begin, cfg.row_end, cfg.dtype = 100, 50, 20, 60, L.F64
    assert lib.fdtd_local_rows_halo(C.byref(cfg), 1) == lib.fdtd_local_rows(C.byref(cfg)) == 40 + 1 + 1
    assert lib.fdtd_local_rows_halo(C.byref(cfg), 7) == 40 + 7 + 7 + 1
    cfg.row_begin, cfg.row_end = 0, 60
    assert lib.fdtd_local_rows_halo(C.byref(cfg), 7) == 60 + 7 + 1
    cfg.row_begin, cfg.row_end = 60, 100
    assert lib.fdtd_local_rows_halo(C.byref(cfg), 7) == 40 + 7 + 1          # the last slab keeps the ex[rows] row
    assert lib.fdtd_set_halo_depth(None, 3) == -1 and lib.fdtd_halo_blocks(None, None, None, None, None, None) == -1
