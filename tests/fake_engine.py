"""TEST DOUBLE -- an `Engine` look-alike that steps with the numpy oracle, so that the host logic of the drop-in
`Solver` (engine life-cycle, snapshot/probe plumbing, update() paths, error behaviour) can be exercised on a box
without a GPU.  Lives in tests/ and is only ever monkeypatched in by tests; the product has no such path."""
import numpy as np

from oracle import restated as R
from fdtd_em_solver_b200.engine import build_source_tables


class FakeEngine:
    created = 0

    def __init__(self, rows, cols=None, *, courant, one_minus_courant=None, dtype="float64", device=0,
                 reflect=(False, False, False, False), row_begin=0, row_end=None, kernel="stream",
                 tile_rows=0, vec=0, block_warps=0, temporal=1, resident=None):
        FakeEngine.created += 1
        cols = rows if cols is None else cols
        assert one_minus_courant is None or one_minus_courant == (1 - courant)
        self.rows, self.cols = rows, cols
        self.f = R.Fields.zeros(rows, cols)
        self.setup = R.Setup(eps_coef=None, mu_coef=None, courant=courant, reflect_walls=tuple(bool(b) for b in reflect))
        self.first_bad = None
        self.n_steps = 0
        self.cells = []
        self.records = []
        self.times = None
        self.closed = False
        self.views = []

    def upload_coefficients(self, eps_coef, mu_coef):
        self.setup.eps_coef, self.setup.mu_coef = np.array(eps_coef), np.array(mu_coef)

    def paint_coefficients(self, shapes, dt_over_ds):
        raise AssertionError("device painting is not part of the CPU double")

    def set_sources(self, pulses, times):
        _, _, _, _, _, self.first_bad = build_source_tables(pulses, times)       # the product's own table logic
        self.times = np.asarray(times, dtype=np.float64)
        self.setup.pulses = [R.PulseSpec(p.get_type(), p.get_direction(), p.get_row(), p.get_col(), p.get_start_time(),
                                         p.get_end_time(), p._magnitude_vector) for p in pulses]

    def set_reflectors(self, rects):
        self.setup.reflectors = [list(map(int, r)) for r in rects]

    def set_probes(self, cells):
        self.cells = [tuple(map(int, c)) for c in cells]
        self.records = []

    def set_fields(self, h=None, ex=None, ey=None):
        z = R.Fields.zeros(self.rows, self.cols)
        self.f.h = np.array(h) if h is not None else z.h
        self.f.ex = np.array(ex) if ex is not None else z.ex
        self.f.ey = np.array(ey) if ey is not None else z.ey

    def zero_fields(self):
        self.set_fields()

    def get_fields(self, which="hxy", out=None):
        return self.f.h.copy(), self.f.ex.copy(), self.f.ey.copy()

    def _one(self, n):
        if self.first_bad is not None and n >= self.first_bad:
            raise IndexError("index %d is out of bounds for the pulse magnitude table" % n)
        self.f.step = n
        R.leapfrog(self.f, self.setup, self.times[n])
        self.n_steps += 1
        if self.cells:
            self.records.append([(self.f.h[i][j], self.f.ex[i][j], self.f.ey[i][j]) for i, j in self.cells])

    def step(self, n):
        self._one(int(n))

    def run(self, first_step, n_steps, snapshot_every=0, total_calls=None, alloc=None):
        saved, out = [], None
        if alloc is not None and snapshot_every:           # like the real engine: the stack exists before the run starts
            k, a, b = int(snapshot_every), int(first_step), int(first_step) + int(n_steps)
            if b // k - a // k:
                out = alloc((b // k - a // k, self.rows, self.cols))
        for n in range(int(first_step), int(first_step) + int(n_steps)):
            self._one(n)                                   # the next call's source write mutates saved[-1]: App. A-Q8
            if snapshot_every and (n + 1) % snapshot_every == 0:
                saved.append(self.f.h)
        if not saved:
            return None
        if out is None:
            return np.array(saved)
        out[...] = saved
        return out

    def run_streamed(self, eps_coef, mu_coef, first_step, n_steps, zero_fields=True, want_h=True, h_out=None):
        self.upload_coefficients(eps_coef, mu_coef)
        self.streamed_runs = getattr(self, "streamed_runs", 0) + 1
        if zero_fields:
            self.set_fields()
        self.run(first_step, n_steps)
        return self.f.h.copy() if want_h else None

    def run_enqueue(self, first_step, n_steps):
        first_step, n_steps = int(first_step), int(n_steps)
        if self.first_bad is not None and first_step + n_steps > self.first_bad:
            if first_step >= self.first_bad:
                raise IndexError("index %d is out of bounds for the pulse magnitude table" % self.first_bad)
            n_steps = self.first_bad - first_step
        for n in range(first_step, first_step + n_steps):
            self._one(n)
        return n_steps

    def view_begin(self, row_stride=1, col_stride=1):
        assert len(self.views) < 2, "at most two pictures in flight"
        self.views.append(self.f.h[::row_stride, ::col_stride].copy())

    def view_end(self):
        return self.views.pop(0)

    def probes(self):
        return np.array(self.records, dtype=np.float64).reshape(len(self.records), len(self.cells), 3)

    def stats(self):
        return dict(steps=self.n_steps, cell_updates=self.n_steps * self.rows * self.cols, kernel_launches=0,
                    halo_bytes_sent=0, last_run_ms=0.0)

    def sync(self):
        pass

    def close(self):
        self.closed = True


class FakeSlabEngine:
    """TEST DOUBLE for a SHARDED run: one y-slab per rank, stepped with the oracle's slab scheme (oracle/slabs.py) and
    the product's halo plan over torch.distributed (slab.exchange, gloo) -- so a multi-process CPU test of the Solver's
    sharding logic really computes its slab from exchanged halo rows, like the CUDA engine does over NCCL."""
    UID = bytes(range(128))

    def __init__(self, rows, cols=None, *, courant, one_minus_courant=None, dtype="float64", device=0,
                 reflect=(False, False, False, False), row_begin=0, row_end=None, kernel="stream",
                 tile_rows=0, vec=0, block_warps=0, temporal=1, resident=None):
        cols = rows if cols is None else cols
        self.rows, self.cols = rows, cols
        self.row_begin, self.row_end = row_begin, rows if row_end is None else row_end
        assert resident is False, "slabs are never resident"
        self.setup = R.Setup(eps_coef=None, mu_coef=None, courant=courant, reflect_walls=tuple(bool(b) for b in reflect))
        self.fl = self.sl = self.plan = None
        self.first_bad, self.n_steps, self.cells, self.records, self.times = None, 0, [], [], None
        self.set_fields()

    @staticmethod
    def comm_unique_id():
        return FakeSlabEngine.UID

    def comm_init(self, rank, nranks, unique_id):
        from fdtd_em_solver_b200 import slab as SL
        assert bytes(unique_id) == self.UID                      # rank 0's id reached every rank
        self.plan = SL.plan_for(rank, nranks, self.rows, self.cols)
        assert (self.plan.row_begin, self.plan.row_end) == (self.row_begin, self.row_end)

    def upload_coefficients(self, eps_coef, mu_coef):
        self.setup.eps_coef, self.setup.mu_coef = np.array(eps_coef), np.array(mu_coef)
        self.sl = None

    def set_sources(self, pulses, times):
        _, _, _, _, _, self.first_bad = build_source_tables(pulses, times)
        self.times = np.asarray(times, dtype=np.float64)
        self.setup.pulses = [R.PulseSpec(p.get_type(), p.get_direction(), p.get_row(), p.get_col(), p.get_start_time(),
                                         p.get_end_time(), p._magnitude_vector) for p in pulses]
        self.sl = None

    def set_reflectors(self, rects):
        self.setup.reflectors = [list(map(int, r)) for r in rects]
        self.sl = None

    def set_probes(self, cells):
        self.cells = [tuple(map(int, c)) for c in cells]
        self.records = []

    def set_fields(self, h=None, ex=None, ey=None):
        from oracle import slabs as OS
        z = R.Fields.zeros(self.rows, self.cols)
        whole = R.Fields(np.array(h) if h is not None else z.h, np.array(ex) if ex is not None else z.ex,
                         np.array(ey) if ey is not None else z.ey)
        self.fl, self.off = OS.make_slab(whole, self.row_begin, self.row_end)      # owned rows + halo rows only

    def zero_fields(self):
        self.set_fields()

    def _own(self):
        ht = self.row_begin - self.off
        return ht, self.row_end - self.row_begin

    def get_fields(self, which="hxy", out=None):
        z = R.Fields.zeros(self.rows, self.cols)                 # rows of other slabs stay zero, as with the real engine
        ht, n = self._own()
        b, e = self.row_begin, self.row_end
        z.h[b:e], z.ey[b:e] = self.fl.h[ht:ht + n], self.fl.ey[ht:ht + n]
        last = 1 if e == self.rows else 0
        z.ex[b:e + last] = self.fl.ex[ht:ht + n + last]
        return z.h, z.ex, z.ey

    def _one(self, n):
        import torch
        import torch.distributed as dist
        from oracle import slabs as OS
        from fdtd_em_solver_b200 import slab as SL
        if self.first_bad is not None and n >= self.first_bad:
            raise IndexError("index %d is out of bounds for the pulse magnitude table" % n)
        if self.sl is None:
            self.sl = OS.local_setup(self.setup, self.off, self.fl.h.shape[0], self.rows)
        self.fl.step = n
        OS.leapfrog_slab(self.fl, self.sl, self.times[n], self.rows, self.off)
        SL.exchange(self.plan, {"h": torch.from_numpy(self.fl.h), "ex": torch.from_numpy(self.fl.ex),
                                "ey": torch.from_numpy(self.fl.ey)}, dist)
        self.n_steps += 1
        if self.cells:
            rec = []
            for i, j in self.cells:
                mine = self.row_begin <= i < self.row_end
                k = i - self.off
                rec.append((self.fl.h[k][j], self.fl.ex[k][j], self.fl.ey[k][j]) if mine else (0.0, 0.0, 0.0))
            self.records.append(rec)

    def step(self, n):
        self._one(int(n))

    def run(self, first_step, n_steps, snapshot_every=0, total_calls=None, alloc=None):
        saved = []
        for n in range(int(first_step), int(first_step) + int(n_steps)):
            self._one(n)                                   # the next call's source write mutates saved[-1]: App. A-Q8
            if snapshot_every and (n + 1) % snapshot_every == 0:
                saved.append(self.fl.h)
        if not saved:
            return None
        ht, rows = self._own()
        stack = np.array([a[ht:ht + rows] for a in saved])
        if alloc is not None:
            out = alloc(stack.shape)
            out[...] = stack
            return out
        return stack

    def probes(self):
        return np.array(self.records, dtype=np.float64).reshape(len(self.records), len(self.cells), 3)

    def stats(self):
        return dict(steps=self.n_steps, cell_updates=self.n_steps * (self.row_end - self.row_begin) * self.cols,
                    kernel_launches=0, halo_bytes_sent=0, last_run_ms=0.0)

    def sync(self):
        pass

    def close(self):
        pass
