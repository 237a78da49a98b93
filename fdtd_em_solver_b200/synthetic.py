"""The C5 synthetic roofline workload (BASELINE.json config 5).

rows x cols Yee grid, heterogeneous eps_r = 1 + 3*u(i,j) from a counter-based splitmix64 (generated on
the device, same bits as oracle.restated.synthetic_coefficients), mu_r = 1, all four walls absorbing,
S = 0.1 with dt/ds taken from Solver(10, 0.1, 4, 1, .) at omega_max = 8e9, one oscillating point source
per slab two rows from the slab interface and one sinusoidal "right" TF/SF line at column 7."""
import math

import numpy as np

from .engine import Engine, C0

DS = 0.011741702542791853                 # math.pi*2*C/(2*8e9)/10, reference solver.py:95-96
DT = DS * 0.1 / C0                        # reference solver.py:97 (stability 0.1)
OMEGA_0, SIGMA_W = 5e9, 1e9


class TablePulse:
    def __init__(self, direction, row, col, start, end, table):
        self.direction, self.row, self.col, self.start, self.end, self.table = direction, row, col, start, end, table


def times(n_calls):
    return np.arange(n_calls) * DT


def pulses(rows, cols, n_calls, slab_edges=(0,)):
    """Same waveform expressions as reference solver.py:541-547."""
    t = times(n_calls)
    sigma_t = 1 / SIGMA_W
    mid = 3 * sigma_t
    osc = np.cos(t * OMEGA_0) * (1 / (sigma_t * np.sqrt(2 * math.pi))) * (np.exp(-((t - mid) ** 2) / (2 * (sigma_t ** 2))))
    out = []
    for edge in slab_edges:
        r = min(max(edge + 2, 2), rows - 3)
        out.append(TablePulse(None, r, min(cols // 3, cols - 3), 0, 0 + mid * 3, osc))
    out.append(TablePulse("right", 0, 7, 0, float(t[-1]) + DT, np.sin(t * OMEGA_0)))
    return out


def make_engine(rows, cols, dtype="float64", n_calls=1024, row_begin=0, row_end=None, slab_edges=None, seed=0,
                device=0, **kw):
    S = (C0 * DT / DS)
    e = Engine(rows, cols, courant=S, one_minus_courant=(1 - S), dtype=dtype, row_begin=row_begin, row_end=row_end,
               device=device, **kw)
    e.fill_synthetic_coefficients(seed, DT / DS)
    t = times(n_calls)
    e.set_sources(pulses(rows, cols, n_calls, slab_edges or (row_begin,)), t)
    return e, t
