"""TEST INFRASTRUCTURE ONLY -- the y-slab scheme of SURVEY.md 8(e) executed with the numpy oracle.

A slab's local arrays (halo row above + owned rows + the ex row below) are stepped as a small grid
of their own with oracle.restated.leapfrog: artificial top/bottom edges are declared "reflective"
(= left untouched), pulses and reflectors are translated to local rows, and pulses that fall outside
stay in the list as inert entries so that the stale `magnitude` local (App. A-Q5) is unchanged.
Owned rows come out bit-identical to the monolithic step; halo rows are then refreshed by the exchange.
"""
import copy

import numpy as np

from . import restated as R


def local_setup(setup, row_off, lrows_h, rows_total):
    """Setup for the sub-grid holding global h rows [row_off, row_off+lrows_h)."""
    s = copy.copy(setup)
    sl = slice(row_off, row_off + lrows_h)
    s.eps_coef, s.mu_coef = setup.eps_coef[sl], setup.mu_coef[sl]
    up, down, left, right = setup.reflect_walls
    s.reflect_walls = (up if row_off == 0 else True, down if row_off + lrows_h == rows_total else True, left, right)
    pulses = []
    for p in setup.pulses:
        q = copy.copy(p)
        q.row = p.row - row_off
        inside = 0 <= q.row < lrows_h
        if p.direction is None and not inside:
            q.direction, q.row = "inert", 0
        elif p.direction == "up" and not (0 <= q.row + 1 < lrows_h):
            q.direction, q.row = "inert", 0
        elif p.direction == "down" and not inside:
            q.direction, q.row = "inert", 0
        elif p.direction in ("left", "right"):
            q.row = 0
        pulses.append(q)
    s.pulses = pulses
    rects = []
    for r0, r1, c0, c1 in setup.reflectors:
        # python slice semantics on the GLOBAL arrays first (ex has rows_total+1 rows), then shift to local rows;
        # the local h/ey arrays end where the global ones do (last slab) or earlier, so numpy clips them the same way
        a0, a1, _ = slice(r0, r1).indices(rows_total + 1)
        rects.append([max(0, a0 - row_off), max(0, a1 - row_off), c0, c1])
    s.reflectors = rects
    return s


def leapfrog_slab(f, s_local, time, rows_total=None, row_off=None):
    """One step of a slab: oracle.restated.leapfrog on the local arrays with the local setup."""
    return R.leapfrog(f, s_local, time)


def make_slab(f_global, row_begin, row_end):
    halo_top = 1 if row_begin > 0 else 0
    off = row_begin - halo_top
    return R.Fields(f_global.h[off:row_end].copy(), f_global.ex[off:row_end + 1].copy(),
                    f_global.ey[off:row_end].copy(), step=f_global.step), off
