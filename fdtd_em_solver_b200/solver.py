"""Drop-in `solver` module: same public surface as the reference's solver.py
(Solver / MaterialArray / Pulse, constants C, MU, EPSILON), with the time-step
path running on the GPU through libfdtd2d.so.

What stays host-side Python (thin, numpy): grid sizing (reference
solver.py:93-105), pulse tables (:480-550), material painting (:352-466),
index conversion (:107-124, :343-345), persistence (:67-91, :303).
What moved to the device: Solver.update (:169-257) and the run loops of
Solver.solve (:288-292, :319-337) -- see engine.py / csrc/.

Extra, optional constructor keywords (defaults keep reference scripts
unchanged): dtype="float64"|"float32"|"float32x" (float32x: fields stored as
fp32 hi + bf16 lo pairs, float64 arithmetic -- the reduced-precision mode that
stays within 1e-5 of the reference on every snapshot; one GPU), device=0,
kernel="stream"|"exact",
material_on_device=False (paint eps/mu on the GPU from the recorded shapes),
temporal_block=None (leapfrog steps per HBM pass in solve(); 2..8 = temporal
blocking, bit-identical results; None = 8 on grids beyond 512 x 512 cells,
which are HBM-bound, 1 on the script-sized ones),
resident=None (True/False: all calls between two snapshots in ONE cooperative
launch for grids that live in L2, i.e. every script of the reference; None =
engine.RESIDENT_DEFAULT / $FDTD_RESIDENT; bit-identical results),
sharded=None (True, or None with FDTD_SHARDED=1 in the environment: under
`torchrun --nproc-per-node N` every rank builds the same Solver and owns one
y-slab of rows on its own GPU, halo rows travel over NCCL inside the library --
see shard.py; False / None without the flag: never, every process is its own
simulation; results bit-identical to one GPU).
"""
import json
import math
import os
import sys

import numpy as np

from . import shard as _shard
from .engine import Engine, empty_hugepages
from .slab import split_rows

C = 2.99 * (10 ** 8)
MU = 1.26 * (10 ** (-6))
EPSILON = 8.85 * (10 ** (-12))

VERBOSE = True
# plot_and_save streams the snapshot stack straight into <name>.npy (np.lib.format.open_memmap) once it is at least
# this big, instead of holding it in RAM and np.save-ing it afterwards (doubleslit: 6387 x 319^2 x 8 B = 5.2 GB)
STREAM_SNAPSHOT_BYTES = 1 << 30


def _say(*a):
    if VERBOSE and os.environ.get("RANK", "0") == "0":       # a sharded (torchrun) run talks through rank 0 only
        print(*a)


def _pyplot():
    import matplotlib.pyplot as plt      # imported lazily: matplotlib is optional on a GPU box
    return plt


class Solver:
    """reference solver.py:14-349."""

    def __init__(self, points_per_wavelength, stability, eps_r_max, mu_r_max, simulation_size, simulation_time=1,
                 dtype="float64", device=0, kernel="stream", material_on_device=False, temporal_block=None,
                 resident=None, sharded=None):
        self.points_per_wavelength = points_per_wavelength
        self.stability = stability
        self.n_max = math.sqrt(eps_r_max * mu_r_max)

        self.omega_max = 0
        self.lambda_min = None
        self.length_x = self.length_y = simulation_size
        self.ds = self.dt = None

        self.steps = None
        self.end_time = simulation_time
        self.size = None
        self.time_vector = None

        self.time = 0
        self.step = 0
        self.h_list = []

        self.material = None
        self.eps_coefficient_arr = None
        self.mu_coefficient_arr = None

        self.reflect = False
        self.reflectors = []
        self.boundaries = [False, False, False, False]     # up, down, left, right; True = reflect
        self.pulses = []
        self.pulses_max = None

        self.save_file_name = None
        self.save_file = False
        self.save_dict_json = dict()
        self.step_frequency = 5
        # NOT in the reference: update() calls per animation frame of plot_and_show.  1 = the
        # reference's one call per frame; n > 1 runs n calls in one device run and shows every n-th field, which keeps
        # the interactive view usable on grids where drawing, not stepping, sets the pace
        self.frame_stride = 1
        # NOT in the reference either: the picture of the realtime view is h[::d, ::d] (a 65536-column grid on a 1000-pixel
        # window needs no more); 1 = every cell.  Only the picture is decimated, never the simulation
        self.view_decimation = 1

        self.recorded_energies = []
        self.energy_vector = []
        self.recorded_fields = []          # record_field(): h series sampled by on-device probes (no snapshots needed)

        # device side
        self._dtype, self._device, self._kernel = dtype, device, kernel
        self._material_on_device = material_on_device
        # leapfrog steps per HBM pass in the run loops (1..8); None = 8 on grids beyond resident mode's reach (> 512 x 512
        # cells: HBM-bound, where temporal blocking pays), 1 on the script-sized ones.  Same bits either way.
        self._temporal_block = None if temporal_block is None else int(temporal_block)
        self._resident = resident                       # None = the engine's default
        self._sharded = sharded                         # None = shard only if FDTD_SHARDED=1 (shard.context)
        self._shard_ctx = None                          # shard.Shard of the current engine, or None
        self._slabs = None                              # [(row_begin, row_end)] per rank of a sharded engine
        self._engine = None
        self._engine_key = None
        self._host = None           # (h, ex, ey) host copies valid for the current step, or None
        self._pending = None        # host fields to upload before the next step
        self._pending_is_zero = False
        self._host_h = None         # h alone, when a streamed run brought it back with its last calls
        self._probe_seen = 0

    # ------------------------------------------------------------------ sizing
    def update_constants(self):
        """reference solver.py:93-105; also re-zeroes the fields."""
        self.lambda_min = math.pi * 2 * C / (self.n_max * self.omega_max)
        self.ds = self.lambda_min / self.points_per_wavelength
        self.dt = min(self.ds * self.stability / C, math.pi / self.omega_max)
        self.size = int(self.length_x / self.ds)
        n = self.size
        self._pending = (np.zeros((n, n)), np.zeros((n + 1, n)), np.zeros((n, n + 1)))
        self._pending_is_zero = True
        self._host = self._pending
        self.steps = int(self.end_time / self.dt)

    def convert(self, si_unit):
        return int((si_unit / self.length_x) * self.size)

    # field views: host copies of the device state (the reference keeps them on the host)
    def _fetch(self):
        """COLLECTIVE in a sharded run: every rank ends up with the whole arrays (slabs gathered over torch.distributed)."""
        if self._host is None:
            got = self._engine.get_fields()
            sh = self._shard_ctx
            if sh is not None:
                b, e = self._slabs[sh.rank]
                last = 1 if e == self.size else 0                    # the last slab also owns ex[size]
                parts = sh.all_gather((got[0][b:e], got[1][b:e + last], got[2][b:e]))
                got = tuple(np.concatenate([part[k] for part in parts]) for k in range(3))
            self._host = got
        return self._host

    h = property(lambda self: None if self.size is None else (self._host_h if self._host_h is not None and self._host is None
                                                               else self._fetch()[0]))
    ex = property(lambda self: None if self.size is None else self._fetch()[1])
    ey = property(lambda self: None if self.size is None else self._fetch()[2])

    def _set_field(self, k, value):
        cur = list(self._fetch())
        cur[k] = np.array(value, dtype=np.float64)
        self._host = self._pending = tuple(cur)
        self._pending_is_zero = False

    h = h.setter(lambda self, v: self._set_field(0, v))
    ex = ex.setter(lambda self, v: self._set_field(1, v))
    ey = ey.setter(lambda self, v: self._set_field(2, v))

    # ------------------------------------------------------------------ scene
    def create_material(self):
        self.material = MaterialArray(self.size, self.length_x, self.length_y, on_device=self._material_on_device)
        return self.material

    def set_reflect_square(self, upper_left, lower_right):
        top, left = (self.convert(v) for v in upper_left)
        bottom, right = (self.convert(v) for v in lower_right)
        self.reflect = True
        self.reflectors.append([top, bottom, left, right])

    def set_reflect_boundaries(self, up=True, down=True, left=True, right=True):
        self.boundaries = [up, down, left, right]

    def _grow(self, omega):
        if omega > self.omega_max:
            self.omega_max = omega
            self.update_constants()

    def _add(self, kind, location, start_time, direction, **kw):
        pulse = Pulse(dt=self.dt, location=[self.convert(v) for v in location], start_time=start_time, type=kind,
                      sim_end_time=self.end_time, direction=direction, real_location=location, **kw)
        self.pulses.append(pulse)
        return pulse

    def add_oscillating_pulse(self, sigma_w, location, omega_0, start_time=0, direction=None):
        self._grow(3 * sigma_w + omega_0)
        return self._add("oscillate", location, start_time, direction, sigma_w=sigma_w, omega_0=omega_0)

    def add_gaussian_pulse(self, sigma_w, location, start_time=0, direction=None):
        self._grow(3 * sigma_w)
        return self._add("gd", location, start_time, direction, sigma_w=sigma_w)

    def add_sinusoidal_pulse(self, omega_0, location, start_time=0, direction=None):
        self._grow(omega_0)
        return self._add("sinusoidal", location, start_time, direction, omega_0=omega_0)

    def ready(self):
        return len(self.pulses) > 0

    def update_reflect(self):
        """reference solver.py:164-167: zero h inside every reflector.  update() applies this on the device as part of
        the step; the method is kept for callers that invoke it by hand between steps (it edits the host view, which
        is uploaded before the next step)."""
        h = np.array(self.h)
        for top, bottom, left, right in self.reflectors:
            h[top:bottom, left:right] = 0
        self.h = h

    def record_energy(self, location):
        row, col = (self.convert(v) for v in location)
        self.recorded_energies.append({"i": row, "j": col, "location": location, "energies": []})

    def record_field(self, location):
        """NOT in the reference: the series a `DataCollector` would extract from the saved snapshot
        stack at `location` (analyser.py:117-124: cell int(x / ds), one sample every `step_frequency` calls), sampled
        by an on-device probe instead -- so an analysis run needs no snapshots at all (`solve(solve_only=True)`).
        Returns the index to pass to `field_collector()` after the run."""
        self.recorded_fields.append({"location": location, "samples": []})
        return len(self.recorded_fields) - 1

    def field_collector(self, index=0):
        """`analyser.DataCollector` over the series of `record_field(...)` number `index`: same `.data`, `.time_vector`,
        `.fft()` and plots as `FileLoader(name).create_data_collector(location)` after a saved run, bit for bit."""
        from .analyser import ProbeCollector
        d = self.recorded_fields[index]
        return ProbeCollector(d["samples"], d["location"], self._run_constants())

    def _field_cells(self):
        return [(int(d["location"][0] / self.ds), int(d["location"][1] / self.ds)) for d in self.recorded_fields]

    def save(self, file_name):
        self.save_file_name = file_name
        self.save_file = True

    def reset(self):
        self.step = 0
        self.update_constants()

    def assign_pulse_max(self):
        self.pulses_max = max(p.pulse_maxima() for p in self.pulses)
        _say('Solver max:', self.pulses_max)

    # ------------------------------------------------------------------ device plumbing
    def _coefficients(self):
        """reference solver.py:278-279 (skipped on the host when the material is painted on the device)."""
        if not self.material.on_host and self.material.ops_complete:
            self.eps_coefficient_arr = self.mu_coefficient_arr = ("device", len(self.material.ops))
            return
        self.eps_coefficient_arr = (self.dt / self.ds) * (1 / self.material.eps_arr)
        self.mu_coefficient_arr = (self.dt / self.ds) * (1 / self.material.mu_arr)

    def _sync_engine(self, times, defer_coefficients=False):
        """(Re)build the device state from the host-side description when it changed.  defer_coefficients: the caller
        sends host coefficient planes itself, pipelined with the run (Engine.run_streamed); returns True if they are
        still to be sent."""
        deferred = False
        courant = (C * self.dt / self.ds)                    # evaluated exactly as solver.py:221 does
        sh = _shard.context(self._sharded)
        temporal = self._temporal_block
        if temporal is None:
            temporal = 8 if (self.size + 1) ** 2 > 512 * 512 and self._kernel == "stream" else 1
        key = (self.size, self._dtype, self._device, self._kernel, tuple(bool(b) for b in self.boundaries), courant,
               temporal, self._resident, None if sh is None else (sh.rank, sh.world))
        if self._engine is None or key != self._engine_key:
            if self._engine is not None:
                if self._pending is None:
                    self._pending = self._fetch()
                self._engine.close()
            if sh is None:
                self._slabs = None
                self._engine = Engine(self.size, courant=courant, one_minus_courant=(1 - courant), dtype=self._dtype,
                                      device=self._device, reflect=self.boundaries, kernel=self._kernel,
                                      temporal=temporal, resident=self._resident)
            else:                   # one y-slab per rank, each on its own GPU; the library exchanges the halo rows itself
                self._slabs = split_rows(self.size, sh.world)
                b, e = self._slabs[sh.rank]
                self._engine = Engine(self.size, courant=courant, one_minus_courant=(1 - courant), dtype=self._dtype,
                                      device=sh.local_rank, reflect=self.boundaries, kernel=self._kernel,
                                      temporal=temporal, resident=False, row_begin=b, row_end=e)
                self._engine.comm_init(sh.rank, sh.world,
                                       sh.broadcast(self._engine.comm_unique_id() if sh.rank == 0 else None))
            self._shard_ctx = sh
            self._engine_key = key
            self._static_key = None
            self._probe_seen = 0
            self._steps_seen = 0
        e = self._engine
        static = (id(self.eps_coefficient_arr), id(self.mu_coefficient_arr),
                  tuple(map(tuple, self.reflectors)) if self.reflect else (),
                  tuple((d["i"], d["j"]) for d in self.recorded_energies), tuple(self._field_cells()),
                  tuple(id(p) for p in self.pulses), np.asarray(times).tobytes())
        if static != self._static_key:
            if isinstance(self.eps_coefficient_arr, tuple):          # painted on the device from the recorded shapes
                e.paint_coefficients(self.material.ops, self.dt / self.ds)
            elif defer_coefficients:
                deferred = True
            else:
                e.upload_coefficients(self.eps_coefficient_arr, self.mu_coefficient_arr)
            e.set_reflectors(self.reflectors if self.reflect else [])
            e.set_probes([(d["i"], d["j"]) for d in self.recorded_energies] + self._field_cells())
            e.set_sources(self.pulses, times)
            self._static_key = static
            self._probe_seen = 0
        if self._pending is not None:
            if self._pending_is_zero:                   # fresh fields (update_constants): cleared on the device, not sent
                e.zero_fields()
            else:
                e.set_fields(*self._pending)
            self._pending = None
            self._pending_is_zero = False
        return deferred

    def _collect_probes(self):
        """Energy formula of reference solver.py:251-255, evaluated on the host from raw device probes; and the h
        samples of record_field()."""
        if not (self.recorded_energies or self.recorded_fields):
            return
        raw = self._engine.probes()
        sh = self._shard_ctx
        if sh is not None:          # a probe cell is written by the rank that owns its row; the others hold zeros there
            cells = [(d["i"], d["j"]) for d in self.recorded_energies] + self._field_cells()
            theirs = sh.all_gather(raw)
            raw = raw.copy()
            for k, (i, _) in enumerate(cells):
                raw[:, k, :] = theirs[_shard.owner_of_row(self._slabs, i)][:, k, :]
        new = raw[self._probe_seen:]
        cells = [self.material.cell(d["i"], d["j"]) for d in self.recorded_energies]
        # scalar arithmetic on purpose: np.float64 ** 2 is libm pow(), which is NOT always the correctly rounded x*x
        # that a vectorised `array ** 2` computes -- the reference's energies are the scalar kind (solver.py:251-255)
        for k, d in enumerate(self.recorded_energies):
            eps_ij, mu_ij = cells[k]
            out = d["energies"]
            for hv, exv, eyv in new[:, k]:                           # iterating float64 rows yields np.float64 scalars
                out.append(0.5 * ((eps_ij * (exv ** 2 + eyv ** 2) + mu_ij * hv ** 2)))
        if self.recorded_fields and len(new):
            # record r of `new` was written by the call that left self.step at `calls[r]`; plot_and_save keeps the
            # field after every call with step % step_frequency == 0 (solver.py:334-337)
            calls = np.arange(self.step - len(new) + 1, self.step + 1)
            keep = np.nonzero(calls % self.step_frequency == 0)[0]
            ne = len(self.recorded_energies)
            for k, (cell, d) in enumerate(zip(self._field_cells(), self.recorded_fields)):
                for r in keep:
                    aliased = self._aliased_source_value(cell, int(calls[r]))
                    d["samples"].append(new[r, ne + k, 0] if aliased is None else aliased)
        self._probe_seen = len(raw)

    def _aliased_source_value(self, cell, next_step):
        """App. A-Q8: h_list holds REFERENCES to self.h, so the NEXT call's hard point-source write (solver.py:180)
        lands in the snapshot taken before it.  Value a saved snapshot shows at `cell` instead of the computed h, or
        None.  (Same rule as the device's snapshot pack: only if a next call exists.)"""
        times = self.time_vector
        if times is None or next_step >= len(times):
            return None
        value = None
        for p in self.pulses:
            if p.get_direction() is None and (p.get_row(), p.get_col()) == tuple(cell) \
                    and p.get_start_time() < times[next_step] < p.get_end_time() \
                    and next_step < len(p._magnitude_vector):
                value = p.magnitude(next_step)
                if self._dtype == "float32":
                    value = np.float64(np.float32(value))            # the array it is written into is float32
        return value

    def _run_constants(self):
        """The scalar run constants save_to_json writes (reference solver.py:68-75)."""
        return {key: getattr(self, key) for key in ('dt', 'ds', 'length_x', 'length_y', 'stability', 'omega_max',
                                                     'end_time', 'step_frequency')}

    def _npy_path(self):
        name = self.save_file_name
        return name if name.endswith(".npy") else name + ".npy"          # np.save's rule

    def _snapshot_alloc(self, shape):
        """Where the snapshot stack of a run lives: RAM, or the .npy file itself for big stacks."""
        nbytes = 8 * int(np.prod(shape))
        if self._shard_ctx is None and self.save_file and not self.h_list and nbytes >= STREAM_SNAPSHOT_BYTES:
            self._streamed = self._npy_path()
            return np.lib.format.open_memmap(self._streamed, mode="w+", dtype=np.float64, shape=tuple(shape))
        return empty_hugepages(shape)

    _streamed = None

    def _run_device(self, times, snapshot_every=0, n_calls=None):
        """The loops of solver.py:288-292 / :331-337 for calls self.step .. len(times)-1 (or the next n_calls of them)."""
        # a run without snapshots on a grid beyond resident mode's reach: coefficient upload, calls and the read of h are one
        # pipelined library call (Engine.run_streamed) instead of three
        streamed = (not snapshot_every and n_calls is None and _shard.context(self._sharded) is None
                    and (self.size + 1) ** 2 > 512 * 512 and self._dtype in ("float64", np.float64))
        deferred = self._sync_engine(times, defer_coefficients=streamed)
        first = self.step
        n = len(times) - first
        if n_calls is not None:
            n = max(0, min(n, int(n_calls)))
        self._host = None
        self._host_h = None
        def account(collect):
            done = self._engine.stats()["steps"]
            self.step = first + (done - self._steps_seen)
            self._steps_seen = done
            if collect:
                self._collect_probes()

        try:
            if deferred:
                stack = None
                try:
                    self._host_h = self._engine.run_streamed(self.eps_coefficient_arr, self.mu_coefficient_arr, first, n,
                                                             zero_fields=False)
                except BaseException:
                    self._static_key = None             # the coefficient planes may not have arrived: send them again
                    raise
            else:
                stack = self._engine.run(first, n, snapshot_every, total_calls=len(times),
                                         alloc=self._snapshot_alloc if snapshot_every else None)
        except IndexError:
            # a pulse outliving its table (App. A-Q10): found on the host from the same tables on every rank, so the
            # collective probe gather of a sharded run is still matched; the records up to the failing call are kept
            account(True)
            raise
        except BaseException:
            # anything else may have hit ONE rank only: no collective on the way out (the other ranks would never match it)
            account(self._shard_ctx is None)
            raise
        account(True)
        return self._whole_stack(stack)

    def _whole_stack(self, stack):
        """Sharded runs: the engine returned this rank's rows of every snapshot.  COLLECTIVE.  A fresh saving run puts
        the slabs straight into <name>.npy (rank 0 creates the file, every rank writes its rows and then maps the whole
        stack -- one node, one page cache); otherwise the slabs are gathered in RAM."""
        sh = self._shard_ctx
        if sh is None or stack is None:
            return stack
        b, e = self._slabs[sh.rank]
        if self.save_file and not self.h_list:
            path, shape = self._npy_path(), (stack.shape[0], self.size, stack.shape[2])
            if sh.rank == 0:
                np.lib.format.open_memmap(path, mode="w+", dtype=np.float64, shape=shape).flush()
            sh.barrier()
            whole = np.lib.format.open_memmap(path, mode="r+")
            whole[:, b:e, :] = stack
            whole.flush()
            sh.barrier()
            self._streamed = path
            return whole
        return np.concatenate(sh.all_gather(stack), axis=1)

    def update(self, time):
        """One leapfrog step (reference solver.py:169-257); returns the new h as a host array."""
        if self.eps_coefficient_arr is None:
            self._coefficients()
        table_times = self.time_vector if self.time_vector is not None else np.arange(0, self.end_time, self.dt)
        on_axis = self.step < len(table_times) and table_times[self.step] == time
        if on_axis:
            self._sync_engine(table_times)
            self._engine.step(self.step)
        else:
            # off-axis call: gate on the caller's `time`, index tables with self.step (solver.py:177-178)
            times = np.full(self.step + 1, float(time))
            for p in self.pulses:
                if p.get_start_time() < time < p.get_end_time():
                    p.magnitude(self.step)          # IndexError exactly where the reference raises it
            self._static_key = None
            self._sync_engine(times)
            self._engine.step(self.step)
            self._static_key = None
        self._steps_seen = self._engine.stats()["steps"]
        self.step += 1
        self._host = None
        self._collect_probes()
        return self.h

    # ------------------------------------------------------------------ run loops
    def solve(self, realtime=True, step_frequency=5, solve_only=False):
        if not self.ready():
            raise Exception('No pulse added')

        _say("Matrix size: {}".format(self.size))
        _say("Number of time steps:", self.steps)

        self._coefficients()
        self.assign_pulse_max()
        self.time_vector = np.arange(0, self.end_time, self.dt)
        self.step_frequency = step_frequency

        if solve_only:
            self._loop(0)
            return

        if realtime and not self.save_file:
            self.plot_and_show()
        else:
            if not self.save_file:
                raise Exception("Must create a save file name for non-realtime simulation")
            self.plot_and_save(step_frequency=step_frequency)
            _say("Saving to npy file...")
            writer = self._shard_ctx is None or self._shard_ctx.rank == 0      # sharded: every rank holds the same data
            if self._streamed is None and writer:
                np.save(self.save_file_name, self._stack if self._stack is not None else self.h_list)
            self._streamed = None             # (a streamed stack already IS the file; the page cache keeps it coherent)
            _say("Saving to txt file...")
            if writer:
                self.save_to_json()
            if self._shard_ctx is not None:
                self._shard_ctx.barrier()     # the files exist on return, on every rank
            _say("Complete!")

    def _loop(self, snapshot_every):
        if self.step == 0:
            stack = self._run_device(self.time_vector, snapshot_every)
        else:                       # solve() called again without reset(): replay the reference literally
            stack = None
            frames = []
            for t in self.time_vector:
                self.update(t)
                if snapshot_every and self.step % snapshot_every == 0:
                    frames.append(self.h)
            if frames:
                stack = np.array(frames)
        if self.steps and VERBOSE and os.environ.get("RANK", "0") == "0":
            sys.stdout.write("\r%d%%" % (100 * self.step // self.steps))
            sys.stdout.flush()
        return stack

    def plot_and_save(self, step_frequency):
        """reference solver.py:331-337: headless loop + snapshot capture every `step_frequency` calls."""
        fresh = not self.h_list
        self._streamed = None
        try:
            stack = self._loop(step_frequency if self.save_file else 0)
        except BaseException:
            if self._streamed is not None:         # the reference writes no file when the loop raises
                path, self._streamed = self._streamed, None
                if os.path.exists(path):
                    os.remove(path)
            raise
        if stack is not None:
            self.h_list.extend(stack[k] for k in range(len(stack)))
        self._stack = stack if fresh else None

    _stack = None

    def plot_and_show(self):
        """reference solver.py:309-329: one update per animation frame.  With frame_stride > 1 or view_decimation > 1
         a frame is a batch of frame_stride calls, and the device runs one batch ahead of the picture:
        while matplotlib draws the picture of batch k, batch k+1 is stepping and the (decimated) picture of batch k has
        already left for the host on the copy stream."""
        if _shard.context(self._sharded) is not None:
            raise Exception("realtime plotting is not available in a sharded run: save() and solve(realtime=False)")
        plt = _pyplot()
        import matplotlib.animation as animation
        stride, dec = max(1, int(self.frame_stride)), max(1, int(self.view_decimation))
        fig, ax1 = plt.subplots(figsize=(6, 6))
        im = ax1.imshow(self.h if dec == 1 else self.h[::dec, ::dec], animated=True, extent=[0, self.length_x, self.length_y, 0],
                        vmax=self.pulses_max, vmin=-self.pulses_max, aspect='auto', cmap='seismic')
        ax1.set_aspect('equal', 'box')
        ax1.xaxis.set_ticks_position('top')
        ax1.xaxis.set_label_position('top')
        fig.colorbar(im, ax=ax1)

        if stride == 1 and dec == 1:
            def animate(time):
                if self.step % self.step_frequency == 0:
                    fig.suptitle("Time Step = {}".format(self.step))
                im.set_array(self.update(time))
                return im,
        else:
            self._sync_engine(self.time_vector)
            pipe = self._pipeline = {"queued": self.step, "batches": [], "error": None, "first": self.step}

            def enqueue():
                n = min(stride, len(self.time_vector) - pipe["queued"])
                if n <= 0 or pipe["error"] is not None:
                    return
                try:
                    done = self._engine.run_enqueue(pipe["queued"], n)
                except IndexError as exc:           # a pulse outliving its table (App. A-Q10): raised at ITS frame
                    pipe["error"] = exc
                    return
                self._host = None
                self._engine.view_begin(dec, dec)
                pipe["queued"] += done
                pipe["batches"].append(done)

            def animate(time):
                fig.suptitle("Time Step = {}".format(self.step))
                if not pipe["batches"]:
                    enqueue()                       # the first frame's own batch
                if not pipe["batches"]:
                    if pipe["error"] is not None:
                        raise pipe["error"]
                    return im,
                enqueue()                           # the next frame's batch steps while this frame's picture is drawn
                im.set_array(self._engine.view_end())
                self.step += pipe["batches"].pop(0)
                return im,

        try:
            self._anim = animation.FuncAnimation(fig, animate, frames=self.time_vector[::stride], interval=1, blit=False,
                                                 repeat=False)
            plt.show()
        finally:
            self._finish_pipeline()

    _pipeline = None

    def _finish_pipeline(self):
        """After the realtime view: pictures still in flight are fetched (the window was closed early), the step counter
        follows the device, and the probe records of the whole view are collected."""
        pipe, self._pipeline = self._pipeline, None
        if pipe is None:
            return
        while pipe["batches"]:
            self._engine.view_end()
            self.step += pipe["batches"].pop(0)
        self._engine.sync()
        self._steps_seen = self._engine.stats()["steps"]
        self._host = None
        self._collect_probes()

    # ------------------------------------------------------------------ persistence
    def save_to_json(self):
        """reference solver.py:67-91 -- same keys, same order, same file name."""
        d = self.save_dict_json
        for key in ('dt', 'ds', 'length_x', 'length_y', 'stability', 'omega_max', 'end_time', 'step_frequency'):
            d[key] = getattr(self, key)
        d['material'] = self.material.eps_arr.tolist()
        d['pulses'] = [{'sigma_w': 0 if p.sigma_w is None else p.sigma_w, 'omega_0': p.omega_0,
                        'omega_max': p.get_omega_max(), 'location': p.real_location, 'type': p.get_type()}
                       for p in self.pulses]
        with open(self.save_file_name + '.txt', 'w') as outfile:
            json.dump(d, outfile)


class MaterialArray:
    """reference solver.py:352-475: painting of absolute eps/mu per cell.

    Every painter also records its shape (in grid indices) in `self.ops`; with `on_device=True` no host arrays are
    allocated and the solver paints the coefficient planes on the GPU from that list (fdtd_paint_coeffs) -- the route
    for grids whose eps/mu arrays do not fit in host RAM.  `eps_arr` / `mu_arr` then materialise lazily on access."""

    def __init__(self, size, real_size_x, real_size_y, eps_array=None, on_device=False):
        self.size = size
        self.length_x = real_size_x
        self.length_y = real_size_y
        self.ops = []                      # [(kind, (ints...), eps_abs, mu_abs)] in painting order
        self.ops_complete = eps_array is None
        self._eps = eps_array
        self._mu = None
        if not (on_device and eps_array is None):
            self._allocate()

    def _allocate(self):
        if self._eps is None:
            self._eps = np.full((self.size, self.size), 1.0) * EPSILON
        self._mu = np.full((self.size, self.size), 1.0) * MU

    def _host(self):
        """Materialise the host arrays by replaying the recorded shapes (device-painted materials only)."""
        if self._mu is None:
            self._allocate()
            ops, self.ops = self.ops, []
            for kind, a, eps, mu in ops:
                self._apply(kind, a, eps / EPSILON, mu / MU, eps, mu)
            self.ops = ops
        return self

    @property
    def eps_arr(self):
        return self._host()._eps

    @eps_arr.setter
    def eps_arr(self, value):
        self._host()
        self._eps = value
        self.ops_complete = False           # arrays edited by hand: the recorded shapes no longer describe them

    @property
    def mu_arr(self):
        return self._host()._mu

    @mu_arr.setter
    def mu_arr(self, value):
        self._host()
        self._mu = value
        self.ops_complete = False

    @property
    def on_host(self):
        return self._mu is not None

    def cell(self, i, j):
        """(eps, mu) of one cell without materialising the arrays (used by the energy probes)."""
        if self.on_host:
            return self._eps[i][j], self._mu[i][j]
        n = self.size
        i, j = (i + n if i < 0 else i), (j + n if j < 0 else j)
        eps, mu = np.float64(1.0 * EPSILON), np.float64(1.0 * MU)
        for kind, a, e, m in self.ops:
            if kind == 0:
                hit = a[0] <= i < a[1] and a[2] <= j < a[3]
            elif kind == 1:
                x, y = i - a[0], j - a[1]
                hit = a[2] ** 2 > x * x + y * y > (a[2] - a[3]) ** 2 and x < 0 and y > 0
            else:
                ci, cj, rad, disp, _ = a
                hit = False
                for ii in (i - ci, i - n - ci):
                    for cc in (j, j - n):
                        for jj, lo, hi in ((cc - cj - disp, -rad, -disp), (cc - cj + disp, disp, rad)):
                            if -rad <= ii < rad and lo <= jj < hi and ii * ii + jj * jj < rad * rad:
                                hit = True
            if hit:
                eps, mu = np.float64(e), np.float64(m)
        return eps, mu

    def convert(self, si_unit):
        if si_unit == self.length_x:
            return self.size
        return int((si_unit / self.length_x) * self.size)

    def _record(self, kind, a, epsilon_rel, mu_rel):
        eps, mu = epsilon_rel * EPSILON, mu_rel * MU
        self.ops.append((kind, tuple(int(v) for v in a), eps, mu))
        if self.on_host:
            self._apply(kind, a, epsilon_rel, mu_rel, eps, mu)

    def _apply(self, kind, a, epsilon_rel, mu_rel, eps, mu):
        n = self.size
        if kind == 0:
            rows, cols = slice(a[0], a[1]), slice(a[2], a[3])
        elif kind == 1:
            di = np.arange(-a[0], n - a[0])[:, None]
            dj = np.arange(-a[1], n - a[1])[None, :]
            d2 = di ** 2 + dj ** 2
            mask = (a[2] ** 2 > d2) & (d2 > (a[2] - a[3]) ** 2) & (di < 0) & (dj > 0)
            self._eps[mask] = eps
            self._mu[mask] = mu
            return
        else:
            ci, cj, radius, displacer, _ = a
            i = np.arange(-radius, radius)[:, None]
            for js, shift in ((np.arange(-radius, -displacer), displacer), (np.arange(displacer, radius), -displacer)):
                j = js[None, :]
                rr, cc = np.broadcast_arrays(ci + i, cj + j + shift)
                keep = (i ** 2 + j ** 2 < radius ** 2) & (rr >= -n) & (rr < n) & (cc >= -n) & (cc < n)
                self._eps[rr[keep] % n, cc[keep] % n] = eps
                self._mu[rr[keep] % n, cc[keep] % n] = mu
            return
        self._eps[rows, cols] = eps
        self._mu[rows, cols] = mu

    def set_material_rect(self, upper_left, lower_right, epsilon_rel, mu_rel=1, convert=True):
        """Rows end-exclusive, columns end-INCLUSIVE (reference solver.py:374-375)."""
        if convert:
            upper_left = [self.convert(v) for v in upper_left]
            lower_right = [self.convert(v) for v in lower_right]
        r0, r1, _ = slice(upper_left[0], lower_right[0]).indices(self.size)      # python slice clipping
        c0, c1, _ = slice(upper_left[1], lower_right[1] + 1).indices(self.size)
        self._record(0, (r0, max(r0, r1), c0, max(c0, c1)), epsilon_rel, mu_rel)

    def set_material_convex(self, center: tuple, radius, thickness, epsilon_rel, mu_rel=1, convert=True):
        """Two circular caps facing each other (reference solver.py:377-409), vectorised.

        The reference indexes cell by cell inside try/except IndexError, so indices
        in [-size, size) are legal (negative ones wrap) and anything else is skipped."""
        if convert:
            displacer = self.convert(radius - (thickness / 2))
            ci, cj = self.convert(center[0]), self.convert(center[1])
            radius = self.convert(radius)
        else:
            displacer = radius - thickness // 2
            ci, cj = center[0], center[1]
        self._record(2, (ci, cj, radius, displacer, self.size), epsilon_rel, mu_rel)

    def set_fixed_length_waveguide(self, upper_left, waveguide_length, curved_portion_ratio, thickness, epsilon_rel,
                                   mu_rel=1):
        straight = (waveguide_length * (1 - curved_portion_ratio)) / 2
        bend = curved_portion_ratio * waveguide_length / (math.pi / 2)
        self.set_waveguide(upper_left, thickness, straight, bend + thickness / 2, epsilon_rel, mu_rel=mu_rel,
                           convert=True)

    def set_quarter_circle(self, center, radius, epsilon_rel, mu_rel=1, thickness=0, convert=True):
        """Upper-right quarter annulus (reference solver.py:418-432)."""
        ci, cj = center[0], center[1]
        if convert:
            ci, cj = self.convert(ci), self.convert(cj)
            radius = self.convert(radius)
            thickness = self.convert(thickness)
        print("latest one   ")
        self._record(1, (ci, cj, radius, thickness), epsilon_rel, mu_rel)

    def set_waveguide(self, start_point, thickness, rect_length, radius, epsilon_rel, mu_rel=1, convert=True):
        """Straight - quarter bend - straight (reference solver.py:446-466)."""
        i0, j0 = start_point[0], start_point[1]
        if convert:
            i0, j0, thickness, rect_length, radius = (self.convert(v) for v in (i0, j0, thickness, rect_length, radius))
        self.set_material_rect((i0, j0), (i0 + thickness, j0 + rect_length), epsilon_rel, mu_rel=mu_rel, convert=False)
        bend_center = (i0 + radius, j0 + rect_length)
        self.set_quarter_circle(bend_center, radius, epsilon_rel, thickness=thickness, mu_rel=mu_rel, convert=False)
        leg = (bend_center[0], bend_center[1] + radius - thickness)
        self.set_material_rect(leg, (leg[0] + rect_length, leg[1] + thickness), epsilon_rel, mu_rel=mu_rel,
                               convert=False)

    def plot(self):
        plt = _pyplot()
        fig, ax = plt.subplots()
        image = ax.imshow(self.eps_arr, extent=([0, self.length_x, self.length_y, 0]), vmax=(np.max(self.eps_arr)),
                          vmin=0, aspect='auto')
        ax.set_aspect('equal', 'box')
        ax.set_title('Material (epsilon relative)')
        fig.colorbar(image, ax=ax)
        plt.show()


def _envelope(t, sigma_t, mid):
    return (1 / (sigma_t * np.sqrt(2 * math.pi))) * (np.exp(-((t - mid) ** 2) / (2 * (sigma_t ** 2))))


class Pulse:
    """reference solver.py:478-642: the source waveform, tabulated once on the host."""

    TYPES = ("gd", "oscillate", "sinusoidal")
    DIRECTIONS = ("up", "down", "left", "right", None)

    def __init__(self, dt, start_time, type, sim_end_time, sigma_w=None, location=None, real_location=None, omega_0=0,
                 direction=None, step_frequency=1):
        _say("--- Pulse Info ---")
        self.omega_0 = omega_0
        self.dt = dt
        self._start_time = start_time
        self._sim_end_time = sim_end_time
        self._index_location = location
        self.real_location = real_location
        self._type = type
        self.plot_title = "Oscillating pulse" if type == "oscillate" else "Gaussian derivative"
        self.step_frequency = step_frequency
        self.sigma_w = sigma_w

        if sigma_w is not None:
            self.sigma_t = 1 / sigma_w
            self.pulse_mid = 3 * self.sigma_t
            self._end_time = start_time + self.pulse_mid * 3
            self._omega_max = 3 * sigma_w + omega_0
            _say("Pulse type: {}, sigma_w: {}, omega_0: {}, omega_max: {}".format(type, sigma_w, omega_0,
                                                                                 self._omega_max))
        else:
            self._end_time = sim_end_time
            self._omega_max = omega_0
            _say("Pulse type: {}, omega_0: {}".format(type, omega_0))

        self._direction = direction
        self._end_step = int(1 + (self._end_time // dt))
        self._time_vector = np.arange(0, sim_end_time, dt)
        self._magnitude_vector = None
        self.create_magnitude_vector()
        self._maximum = self.pulse_maxima()
        self.fft_amplitude = None
        self.fft_frequencies = None

        if type not in self.TYPES:
            raise Exception("Pulse type not recognised: {}".format(type))
        if direction not in self.DIRECTIONS:
            raise Exception("Pulse direction not recognised: {}".format(self.get_direction))
        if type == "oscillate" and not omega_0:
            raise Exception("Missing arguments f_0: {}".format(omega_0))

    def magnitude(self, timestep):
        return self._magnitude_vector[timestep]

    def pulse_maxima(self):
        return None if self._magnitude_vector is None else np.max(self._magnitude_vector)

    def create_magnitude_vector(self):
        t = self._time_vector
        if self._type == "sinusoidal":
            self._magnitude_vector = np.sin(t * self.omega_0)
            return
        carrier = None
        if self._type == "gd":
            carrier = -t + self.pulse_mid
        elif self._type == "oscillate":
            carrier = np.cos(t * self.omega_0)
        if carrier is None:                 # unknown type: the constructor raises right after (reference :523-524)
            return
        # association order of reference solver.py:541-545: (carrier * norm) * gaussian, rounded in that order
        norm = 1 / (self.sigma_t * np.sqrt(2 * math.pi))
        self._magnitude_vector = carrier * norm * np.exp(-((t - self.pulse_mid) ** 2) / (2 * (self.sigma_t ** 2)))
        _say("Pulse duration: {} seconds, {} steps".format(t[-1], len(t)))

    # ---- analysis of the source waveform (host, out of the hot path) ----
    def plot_time(self, scaled_plot=True, show=True, title=False):
        plt = _pyplot()
        plt.plot(self._time_vector, self._magnitude_vector)
        if title:
            plt.suptitle(self.plot_title)
        if self._type == "oscillate":
            local = np.arange(0, self._end_time, self.dt)
            plt.plot(local, _envelope(local, self.sigma_t, self.pulse_mid))
        plt.xlabel("Time (seconds)")
        plt.ylabel("Magnitude")
        if scaled_plot:
            plt.xlim(0, self._end_time)
        if show:
            plt.show()

    def plot_frequency(self, show=True, title=False):
        plt = _pyplot()
        plt.xlabel("Frequency (rad/s)")
        plt.ylabel("Amplitude")
        if title:
            plt.suptitle(self.plot_title)
        w = np.linspace(self._omega_max, 1000)
        if self._type == "oscillate":
            plt.plot(w, self.sigma_t * 0.5 * (np.exp((-(w - self.omega_0) ** 2 / (2 * self.sigma_w ** 2))) +
                                              np.exp((-(w + self.omega_0) ** 2 / (2 * self.sigma_w ** 2)))))
        elif self._type == "gd":
            plt.plot(w, np.abs(w) * self.sigma_t * np.exp(-((w ** 2) / (2 * self.sigma_w ** 2))))
        if show:
            plt.show()

    def fft(self):
        samples = self._magnitude_vector[::self.step_frequency]
        padded = np.concatenate((samples, np.zeros(2 ** 14 - len(samples))))
        freqs = 2 * np.pi * np.fft.fftfreq(len(padded), d=self.dt * self.step_frequency)
        keep = np.logical_and(freqs > 0, freqs < self._omega_max)
        self.fft_amplitude = np.abs(np.fft.fft(padded)[keep])
        self.fft_frequencies = freqs[keep]
        _say("Number of frequency points:", len(self.fft_frequencies))

    def plot_frequency_fft(self, show=True, title=False):
        plt = _pyplot()
        if self.fft_amplitude is None:
            self.fft()
        plt.plot(self.fft_frequencies, self.fft_amplitude)
        if title:
            plt.suptitle(self.plot_title)
        if show:
            plt.ylabel("Amplitude")
            plt.xlabel("Frequency (rad/s)")
            plt.show()

    def get_maximum(self): return self._maximum
    def get_start_time(self): return self._start_time
    def get_direction(self): return self._direction
    def get_end_time(self): return self._end_time
    def get_end_step(self): return self._end_step
    def get_omega_max(self): return self._omega_max
    def get_row(self): return self._index_location[0]
    def get_col(self): return self._index_location[1]
    def get_type(self): return self._type


def test():
    """The reference's built-in demo scene (solver.py:645-665): a point pulse next to an eps_r = mu_r = 3 block, shown
    in the realtime view (needs matplotlib and a display, like the reference)."""
    solver = Solver(points_per_wavelength=10, stability=0.1, eps_r_max=4, mu_r_max=3, simulation_size=3,
                    simulation_time=2.5e-8)
    _say(C / 5e9)
    solver.add_oscillating_pulse(1e9, (0.8, 0.4), 5e9)
    solver.create_material().set_material_rect((2, 2), (2.8, 2.8), 3, 3)
    solver.solve(step_frequency=50)
    return solver


if __name__ == '__main__':
    test()
