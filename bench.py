#!/usr/bin/env python
"""bench.py -- BASELINE.json's metric on BASELINE.json's config.

metric  : Gcell-updates/s of the 2D Yee leapfrog step (one cell-update = H, Ex, Ey of one cell
          advanced one step, incl. source / wall handling), fp64 by default (--dtype f32).
workload: the C5 synthetic slab: 8192 rows x 65536 cols per GPU of the
          65536 x 65536 heterogeneous-eps grid, y-slab sharded (weak scaling: 8192*N rows at N GPUs),
          absorbing walls, one point source per slab + one "right" TF/SF line at column 7.
a step  : ONE leapfrog time step over the whole (sharded) grid.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--dtype f64|f32] [--impl reference]

N > 1 is launched by torchrun (one rank per GPU); ranks exchange halo rows over NCCL inside the
C library.  Timing: CUDA events on the library's launch stream around fdtd_run, barrier +
synchronize on both sides, max over ranks.  Inputs (8 arrays x 4 GiB per GPU) exceed L2 (126 MB),
so no explicit L2 flush is needed between steps.

`--impl reference` times the reference's CPU path (numpy; the oracle port oracle/restated.py,
bit-identical to /root/reference/solver.py:169-257 which cannot travel to the GPU box) on the host
cores, each step a bounded sample of the same workload.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

ROWS_PER_GPU = 8192
COLS = 65536
# 5 words read + 3 written; f32x = split field storage: 3 fields x 6 B x 2 + 2 float64 coefficients
BYTES_PER_CELL = {"f64": 64, "f32": 32, "f32x": 52}
METRIC = "Gcell-updates/s (2D Yee leapfrog, C5 synthetic slab)"
CPU_SAMPLE = (512, 65536)           # rows x cols of the CPU arm's sample: a thin slab of the real rows
# the plain K-step kernel the library launches when the mu_coef plane is uniform (FDTD_SKEW=0 keeps the un-skewed pipeline)
UMU_KERNEL = "stepK_lean_umu" if os.environ.get("FDTD_SKEW", "1") == "0" else "stepK_skew_umu"


def measured_peak_gbs():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.proc, self.lines = index, None, []

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._pump, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        sm, mx, reasons, watts = [], [], set(), []
        names = ("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap")
        for line in self.lines:
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 7:
                continue
            try:
                sm.append(float(parts[0])); mx.append(float(parts[1])); watts.append(float(parts[2]))
            except ValueError:
                continue
            for name, flag in zip(names, parts[3:7]):
                if flag.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "power_w_max": max(watts) if watts else None, "samples": len(sm), "reasons": sorted(reasons)}


def _reference_module():
    """The UNMODIFIED reference solver.py, if __graft_entry__.build() found /root/reference and put a copy under the
    git-ignored baseline/_ref/ (it travels to the GPU box with the snapshot; /root/reference itself does not)."""
    path = os.path.join(ROOT, "baseline", "_ref", "solver.py")
    if not os.path.isfile(path):
        return None
    try:
        from oracle import ref_import
        ref_import.install_matplotlib_stub()
        import importlib.util
        spec = importlib.util.spec_from_file_location("_fdtd_reference_solver_bench", path)
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)
        return mod
    except Exception:
        return None


def cpu_baseline(steps_budget_s=15.0):
    """The reference's CPU implementation of the step on a bounded, slab-shaped sample of the C5 workload (CPU_SAMPLE
    rows of the real 65536-column rows), timed on this host.  With baseline/_ref/solver.py present it is the reference's
    own Solver.update() (kind "reference"): a reference Solver object whose arrays, coefficient planes and pulses are
    replaced by the sample's (update() only reads attributes, /root/reference/solver.py:169-257); otherwise the oracle
    port oracle/restated.py, bit-identical to it (kind "port").  Returns (Gcell/s, steps, seconds, kind)."""
    from oracle import restated as R
    from fdtd_em_solver_b200 import synthetic as SY
    rows, cols = CPU_SAMPLE
    epc, muc = R.synthetic_coefficients(rows, cols, SY.DT, SY.DS, seed=0, total_cols=COLS)
    n_calls = 512                                    # upper bound; the time budget below ends the loop (10-30 s of CPU work)
    specs = SY.pulses(rows, cols, n_calls)
    t = SY.times(n_calls)
    ref = _reference_module()
    if ref is not None:
        try:
            import contextlib
            import io
            quiet = contextlib.redirect_stdout(io.StringIO())    # the reference prints pulse info: keep stdout one JSON line
            quiet.__enter__()
            sol = ref.Solver(points_per_wavelength=10, stability=0.1, eps_r_max=4, mu_r_max=1, simulation_size=1,
                             simulation_time=1e-9)
            sol.dt, sol.ds = SY.DT, SY.DS
            sol.h, sol.ex, sol.ey = np.zeros((rows, cols)), np.zeros((rows + 1, cols)), np.zeros((rows, cols + 1))
            sol.eps_coefficient_arr, sol.mu_coefficient_arr = epc, muc
            sol.boundaries = [False, False, False, False]
            sol.pulses = []
            for p in specs:
                q = ref.Pulse(sigma_w=SY.SIGMA_W, dt=SY.DT, location=[p.row, p.col], omega_0=SY.OMEGA_0, start_time=0,
                              type="oscillate", sim_end_time=float(t[-1]) + SY.DT, direction=p.direction,
                              real_location=[0, 0])
                q._magnitude_vector = np.asarray(p.table)                # the C5 recipe's tables and windows
                q._start_time, q._end_time = p.start, p.end
                sol.pulses.append(q)
            quiet.__exit__(None, None, None)
            sol.step = 0
            sol.update(t[0])                         # warm-up
            n, t0 = 0, time.perf_counter()
            while n < n_calls - 1 and (n < 3 or time.perf_counter() - t0 < steps_budget_s):
                sol.update(t[n + 1])
                n += 1
            dt = time.perf_counter() - t0
            return rows * cols * n / dt / 1e9, n, dt, "reference"
        except Exception as exc:                     # an API mismatch must not cost the bench line: fall back to the port
            sys.stdout = sys.__stdout__
            sys.stderr.write("bench.py: live reference unusable (%r), timing the oracle port instead\n" % (exc,))
    pulses = [R.PulseSpec("table", p.direction, p.row, p.col, p.start, p.end, p.table) for p in specs]
    S = (SY.C0 * SY.DT / SY.DS)
    setup = R.Setup(eps_coef=epc, mu_coef=muc, courant=S, pulses=pulses)
    f = R.Fields.zeros(rows, cols)
    R.leapfrog(f, setup, t[0])                       # warm-up
    n, t0 = 0, time.perf_counter()
    while n < n_calls - 1 and (n < 3 or time.perf_counter() - t0 < steps_budget_s):
        R.leapfrog(f, setup, t[n + 1])
        n += 1
    dt = time.perf_counter() - t0
    return rows * cols * n / dt / 1e9, n, dt, "port"


def cpu_baseline_c(budget_s=5.0):
    """Informational: the same sample through the C restatement (oracle/restated.c, bit-identical to the numpy one) with
    every host thread OpenMP gives it -- what a compiled multi-core CPU port of the reference step would reach.
    Returns None when gcc is absent or anything goes wrong: it must never cost the bench line."""
    try:
        from oracle import c_oracle as CO, restated as R
        from fdtd_em_solver_b200 import synthetic as SY
        if not CO.available():
            return None
        CO.lib()
        rows, cols = CPU_SAMPLE
        epc, muc = R.synthetic_coefficients(rows, cols, SY.DT, SY.DS, seed=0, total_cols=COLS)
        n_calls = 2048
        pulses = [R.PulseSpec("table", p.direction, p.row, p.col, p.start, p.end, p.table)
                  for p in SY.pulses(rows, cols, n_calls)]
        setup = R.Setup(eps_coef=epc, mu_coef=muc, courant=(SY.C0 * SY.DT / SY.DS), pulses=pulses)
        f = R.Fields.zeros(rows, cols)
        t = SY.times(n_calls)
        CO.leapfrog(f, setup, t[0])
        n, t0 = 0, time.perf_counter()
        while n < n_calls - 1 and (n < 3 or time.perf_counter() - t0 < budget_s):
            CO.leapfrog(f, setup, t[n + 1])
            n += 1
        dt = time.perf_counter() - t0
        return {"value": rows * cols * n / dt / 1e9, "unit": "Gcell-updates/s", "cores": os.cpu_count(),
                "sample": "%dx%d, %d steps in %.1f s, gcc -O2 + OpenMP" % (rows, cols, n, dt)}
    except Exception:
        return None


def run_reference(args, rank):
    if rank != 0:
        return
    value, n, dt, kind = cpu_baseline(steps_budget_s=max(5.0, min(60.0, 0.2 * (args.steps + args.warmup))))
    sample = "%d x %d slab of the C5 recipe (the real row length), %d steps of %s in %.1f s" % (
        CPU_SAMPLE + (n, "the reference's own Solver.update (baseline/_ref/solver.py)" if kind == "reference"
                      else "the numpy oracle port (bit-identical to the reference)", dt))
    cfg = workload_config(args, 1)
    cfg["workload"] += "; CPU arm: a %d x %d sample of it, throughput scaled per cell" % CPU_SAMPLE
    cfg["sampled_shape"] = list(CPU_SAMPLE)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": "Gcell-updates/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / n, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": value, "unit": "Gcell-updates/s", "cores": 1, "kind": kind, "sample": sample,
                         "host_cores": os.cpu_count(), "threaded_c_port": cpu_baseline_c(),
                         "note": "numpy elementwise ufuncs are single-threaded; oracle/restated.py is bit-identical "
                                 "to reference solver.py:169-257 (tests/test_oracle_vs_reference.py)"},
        "e2e": {"value": value, "unit": "Gcell-updates/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line), flush=True)


def workload_config(args, n_gpus, engine=None):
    fm, kv, hd = getattr(args, "frame_mode", None), getattr(args, "kstep_variant", None), getattr(args, "halo_depth", None)
    frame_mode = engine.frame_mode if engine is not None else (3 if fm is None else int(fm))
    variant = engine.kstep_variant if engine is not None else (1 if kv is None else int(kv))
    halo = engine.halo_depth if engine is not None else (1 if hd is None else int(hd))
    n_arrays = len(engine.buffers) if engine is not None else 8
    kvec = 2 if args.dtype == "f64" else 4
    if args.temporal > 1:
        frame = {0: "masked step_stream frame passes", 3: "stepK_edge for the wall / source / reflector tiles"}[frame_mode]
        umu = engine is not None and variant == 1 and engine.uniform_mu()
        kernel = "%s<%s,%d,%d> (%d steps per HBM pass, %d columns per lane) + %s; tile_rows=%d warps=%d" % (
            (UMU_KERNEL if umu else "stepK_lean") if variant == 1 else "stepK_stream",
            "double" if args.dtype == "f64" else "float", kvec, args.temporal,
            args.temporal, kvec, frame, args.tile_rows, args.warps)
    else:
        kernel = "step_stream vec=%d tile_rows=%d warps=%d" % (args.vec, args.tile_rows, args.warps)
    if args.dtype == "f32x":
        kernel = ("stepK_lean_split<%d> (plain tiles: fp32 hi + bf16 lo planes, float64 arithmetic, %d steps per HBM pass) + "
                  "stepK_edge<double,2,%d,all features,split> for the wall / source / reflector tiles; tile_rows=%d warps=%d"
                  % (args.temporal, args.temporal, args.temporal, args.tile_rows, args.warps))
    return {"workload": "C5 synthetic: %d x %d Yee grid (%d rows per GPU, y-slab sharded), heterogeneous eps_r=1+3u "
                        "(splitmix64), mu_r=1, 4 absorbing walls, 1 point source per slab + 1 TF/SF line at col 7"
                        % (ROWS_PER_GPU * n_gpus, COLS, ROWS_PER_GPU),
            "rows": ROWS_PER_GPU * n_gpus, "cols": COLS, "rows_per_gpu": ROWS_PER_GPU,
            "l2": "inputs_exceed_l2 (%d arrays x %.1f GiB per GPU, no flush needed)"
                  % (n_arrays,
                     ROWS_PER_GPU * COLS * {"f64": 8, "f32": 4, "f32x": 6}[args.dtype] / 2 ** 30),
            "parallelism": ("1 GPU" if n_gpus == 1 else
                            "y-slab x%d, %d halo rows of h/ex/ey each way ONCE per K-step group over NCCL"
                            % (n_gpus, halo) if halo > 1 else
                            "y-slab x%d, 3 halo rows down + 1 up per step over NCCL" % n_gpus),
            "kernel": kernel,
            "temporal_blocking": args.temporal,
            "frame_mode": frame_mode, "kstep_variant": variant, "halo_depth": halo,
            "uniform_mu": bool(engine is not None and engine.uniform_mu())}


def main():
    global ROWS_PER_GPU
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=1000,
                    help="timed leapfrog steps (BASELINE.json's config 5 runs 10 000; 1000 keeps the default run short)")
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--dtype", default="f64", choices=["f64", "f32", "f32x"],
                    help="f32x: fp32 hi + bf16 lo field storage, float64 coefficients and arithmetic (1 GPU)")
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--tile-rows", type=int, default=0, help="rows per warp tile (0 = library default for the chosen --temporal)")
    ap.add_argument("--vec", type=int, default=0)
    ap.add_argument("--warps", type=int, default=0, help="warps per CTA (0 = library default)")
    ap.add_argument("--temporal", type=int, default=8, choices=[1, 2, 3, 4, 5, 6, 7, 8],
                    help="leapfrog steps per HBM pass (>1 = temporal blocking)")
    ap.add_argument("--total-rows", type=int, default=0,
                    help="STRONG scaling: fix the grid at this many rows (e.g. 65536, BASELINE config 5) and give each GPU "
                         "total/N of them; default 0 = weak scaling, %d rows per GPU" % ROWS_PER_GPU)
    ap.add_argument("--frame-mode", type=int, default=None, choices=[0, 3],
                    help="non-plain tiles of a temporal-blocking group: 3 one stepK_edge launch, 0 masked single-step "
                         "passes through scratch generations (default: the library's, 3)")
    ap.add_argument("--kstep-variant", type=int, default=None, choices=[0, 1],
                    help="K-step kernel: 0 stepK_stream, 1 stepK_lean (default: the library's)")
    ap.add_argument("--halo-depth", type=int, default=None,
                    help="y-slabs: rows stored from each neighbouring slab; 1 = one row exchanged every step, K+1 = deep "
                         "halos with ONE exchange per K-step group (default: the library's)")
    ap.add_argument("--snapshot-every", type=int, default=0,
                    help="also time the K steps WITH an h snapshot every N steps read back to the host (device pack -> "
                         "pinned ring -> async D2H on the copy stream -> host stack), SURVEY.md 8(d) 'with and without "
                         "snapshot D2H'; reported as `with_snapshots`")
    ap.add_argument("--repeats", type=int, default=0,
                    help="how often the timed --steps region is repeated back to back; the line reports the median "
                         "(0 = 5 repetitions, 3 for runs longer than 2000 steps)")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-cpu", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    if args.vec == 0:
        args.vec = 4

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.total_rows:
        if args.total_rows % world or args.total_rows // world < 64:
            raise SystemExit("--total-rows must be a multiple of the number of GPUs, at least 64 rows per GPU")
        ROWS_PER_GPU = args.total_rows // world
    if args.impl == "reference":
        run_reference(args, rank)
        return

    import torch
    import torch.distributed as dist
    from fdtd_em_solver_b200 import synthetic as SY
    from fdtd_em_solver_b200.engine import comm_unique_id

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a CUDA device (the product has no CPU path); use --impl reference for the CPU arm")
    torch.cuda.set_device(local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    n_gpus = world
    rows = ROWS_PER_GPU * n_gpus
    dtype = {"f64": "float64", "f32": "float32", "f32x": "float32x"}[args.dtype]
    n_calls = 8 * (args.steps + args.warmup) + 256
    e, _ = SY.make_engine(rows, COLS, dtype=dtype, n_calls=n_calls, row_begin=rank * ROWS_PER_GPU,
                          row_end=(rank + 1) * ROWS_PER_GPU, slab_edges=(rank * ROWS_PER_GPU,), device=local_rank,
                          tile_rows=args.tile_rows, vec=args.vec, block_warps=args.warps, temporal=args.temporal,
                          frame_mode=args.frame_mode, kstep_variant=args.kstep_variant, halo_depth=args.halo_depth)
    if world > 1:
        ident = torch.zeros(128, dtype=torch.uint8, device="cuda")
        if rank == 0:
            ident.copy_(torch.frombuffer(bytearray(comm_unique_id()), dtype=torch.uint8))
        dist.broadcast(ident, 0)
        e.comm_init(rank, world, bytes(ident.cpu().numpy().tobytes()))

    def fence():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def max_over_ranks(ms):
        if world == 1:
            return ms
        t = torch.tensor([ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    tuning = e.tuning()                      # what the library picked where the flags left it open
    args.tile_rows, args.warps = tuning["tile_rows"], tuning["block_warps"]

    # ---- device-resident timing ("value") -------------------------------------------------
    # Warm-up: the W steps asked for, then one untimed pass with the GROUP STRUCTURE of the timed region (full K-step
    # groups, the remainder group, every launch shape and exchange pattern once), so that no first-ever launch of a
    # kernel variant, no lazy module load and no NCCL channel set-up can land inside the timed region -- a 20-step run
    # is three K = 6 groups and one K = 2 group, and a 5-step warm-up alone exercises neither.
    K = max(1, args.temporal)
    step = 0
    e.run(step, args.warmup); step += args.warmup
    structural = args.steps if args.steps <= 4 * K else 2 * K + args.steps % K
    e.run(step, structural); step += structural
    fence()
    # The timed region: exactly --steps leapfrog steps, R times back to back (each bracketed by barrier + synchronize,
    # device-timed by CUDA events on the library's stream, max over ranks); value = the MEDIAN repetition.
    reps = args.repeats if args.repeats > 0 else (5 if args.steps <= 2000 else 3)
    e.kernel_timing(True)
    launches0 = e.stats()["kernel_launches"]
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    wall0 = time.perf_counter()
    rep_ms = []
    for _ in range(reps):
        fence()
        e.run(step, args.steps); step += args.steps       # events on the library stream bracket exactly --steps steps
        fence()
        rep_ms.append(max_over_ranks(e.stats()["last_run_ms"]))
    wall = time.perf_counter() - wall0
    clocks = sampler.stop() if rank == 0 else None
    kt = e.kernel_timing_fetch()
    e.kernel_timing(False)
    st = e.stats()
    ms = float(np.median(rep_ms))
    launches = (st["kernel_launches"] - launches0) // reps
    cells = rows * COLS
    value = cells * args.steps / (ms * 1e-3) / 1e9
    halo_bytes = st["halo_bytes_sent"]

    # ---- the same K steps with field snapshots streamed to the host ------------------------------
    with_snapshots = None
    if args.snapshot_every > 0:
        e.run(step, 2, snapshot_every=2); step += 2      # untimed: allocates the device/pinned snapshot ring (GBs of cudaHostAlloc)
        fence()
        t0 = time.perf_counter()
        stack = e.run(step, args.steps, snapshot_every=args.snapshot_every)       # returns after the last copy has landed
        fence()
        sec = time.perf_counter() - t0
        step += args.steps
        if world > 1:
            t = torch.tensor([sec], dtype=torch.float64, device="cuda")
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            sec = float(t.item())
        n_snap = 0 if stack is None else int(stack.shape[0])
        with_snapshots = {"every": args.snapshot_every, "snapshots": n_snap, "value": cells * args.steps / sec / 1e9,
                          "unit": "Gcell-updates/s", "seconds": sec,
                          "d2h_bytes_per_gpu": n_snap * ROWS_PER_GPU * COLS * 8,
                          "what": "wall clock of fdtd_run incl. the first-touch of the pageable host stack; pack kernel + "
                                  "2-slot pinned ring + copy stream, stepping continues while a snapshot drains"}
        del stack

    # ---- end to end through the host-facing API ("e2e") -----------------------------------
    e2e = None
    if not args.no_e2e:
        owned = ROWS_PER_GPU
        tdt = torch.float64
        deep = e.halo_depth if e.halo_depth > 1 else 0
        lo = (deep or 1) if rank > 0 else 0              # halo rows above the slab travel with the coefficient planes
        hb = deep if rank + 1 < world else 0             # ... and, with deep halos, the ones below it
        # host-resident inputs of this slab: the two coefficient planes, pinned, float64 (what
        # Solver.solve hands over, solver.py:278-279); built once outside the timed region
        stored = owned + lo + hb
        eps_slab = torch.empty((stored + 1, COLS), dtype=tdt, pin_memory=True)
        mu_slab = torch.empty((stored + 1, COLS), dtype=tdt, pin_memory=True)
        eps_slab[:stored].copy_(e.buffers[6][:stored, :COLS].to(tdt))
        mu_slab[:stored].copy_(e.buffers[7][:stored, :COLS].to(tdt))
        h_out = torch.empty((owned, COLS), dtype=tdt, pin_memory=True)
        torch.cuda.synchronize()
        import ctypes as C
        from fdtd_em_solver_b200 import _lib as L
        # the host keeps only this rank's slab (first stored row = halo row above the slab, if any)
        L.check(e.lib, e.handle, e.lib.fdtd_set_host_row_origin(e.handle, rank * ROWS_PER_GPU - lo))

        def host_ptr(t):
            return C.c_void_p(t.data_ptr())

        hp = C.c_void_p(h_out.data_ptr() - lo * COLS * 8)        # h_out row 0 = first OWNED row

        def timed(body):
            fence()
            t0 = time.perf_counter()
            body()
            fence()
            sec = time.perf_counter() - t0
            if world > 1:
                t = torch.tensor([sec], dtype=torch.float64, device="cuda")
                dist.all_reduce(t, op=dist.ReduceOp.MAX)
                sec = float(t.item())
            return sec

        def four_calls():
            L.check(e.lib, e.handle, e.lib.fdtd_upload_coeffs(e.handle, host_ptr(eps_slab), host_ptr(mu_slab), COLS))
            e.zero_fields()
            e.run(0, args.steps)
            L.check(e.lib, e.handle, e.lib.fdtd_get_fields(e.handle, hp, None, None))

        def one_call():
            L.check(e.lib, e.handle, e.lib.fdtd_run_streamed(e.handle, host_ptr(eps_slab), host_ptr(mu_slab), COLS, 1, 0,
                                                             args.steps, hp, COLS))

        sec_seq = timed(four_calls)
        sum_seq = float(h_out[:: max(1, owned // 64)].abs().sum())
        h_out.zero_()
        one_call()                                               # structural warm-up of the banded launches (plans, events)
        h_out.zero_()
        sec = timed(one_call)
        assert float(h_out[:: max(1, owned // 64)].abs().sum()) == sum_seq, "streamed run differs from the four calls"
        h2d = 2 * stored * COLS * 8
        d2h = owned * COLS * 8
        e2e = {"value": cells * args.steps / sec / 1e9, "unit": "Gcell-updates/s",
               "h2d_bytes_per_step": h2d / args.steps, "d2h_bytes_per_step": d2h / args.steps,
               "seconds": sec, "checksum_h": float(h_out[:: max(1, owned // 64)].abs().sum()),
               "what": "fdtd_run_streamed: coefficient planes from pinned host float64 arrays -> zero fields -> all calls -> "
                       "h to pinned host, ONE library call, wall clock incl. all copies, per-GPU bytes; on one GPU the three "
                       "stages are pipelined over row bands (H2D, stepping and D2H overlap), on y-slabs they run one after "
                       "the other (= `sequential`)",
               "sequential": {"value": cells * args.steps / sec_seq / 1e9, "seconds": sec_seq,
                              "what": "fdtd_upload_coeffs + fdtd_zero_fields + fdtd_run + fdtd_get_fields(h), four calls"}}
        launches_e2e = e.stats()["kernel_launches"] - st["kernel_launches"]
    else:
        launches_e2e = 0

    if rank != 0:
        if world > 1:
            dist.barrier()
            dist.destroy_process_group()
        return

    peak, peak_src = measured_peak_gbs()
    bpc = BYTES_PER_CELL[args.dtype]
    tname = "double" if args.dtype == "f64" else "float"
    kvec = 2 if args.dtype == "f64" else 4
    if args.temporal > 1 and kt["launches"] > 0:
        # the dominant kernel's OWN duration: CUDA events around every launch of the plain K-step kernel on the stream it
        # is launched on, inside the timed region (fdtd_kernel_timing); its algorithmic bytes = the cell-updates those
        # launches advanced (plain tiles x K steps, counted by the library from its tile plan) x 64 / 32 B
        launch_ms = kt["total_ms"] / kt["launches"]
        alg_bytes = kt["cell_updates"] / kt["launches"] * bpc
        kname = ("fdtd::%s<%s,%d,%d>: one launch advances its tiles %d steps (algorithmic bytes are counted per step, so "
                 "frac > 1 is the temporal-blocking gain; the DRAM-level view is traffic / dram_frac)"
                 % ((UMU_KERNEL if e.uniform_mu() else "stepK_lean") if e.kstep_variant == 1 else "stepK_stream",
                    tname, kvec, args.temporal, args.temporal))
        if e.kstep_variant == 1 and e.uniform_mu():
            kname += ("; mu_coef is one value everywhere (mu_r = 1, the C5 recipe), found by a device scan: the kernel streams "
                      "7 words per cell and pass instead of 8 -- algorithmic bytes stay SURVEY.md's 64 B per cell-update")
    else:
        launch_ms = ms / args.steps
        alg_bytes = ROWS_PER_GPU * COLS * bpc
        kname = "fdtd::step_stream<%s,%d> (one launch per step per GPU%s)" % (
            tname, args.vec, "; 3 launches with halo overlap" if n_gpus > 1 else "")
    if args.dtype == "f32x" and args.temporal > 1 and kt["launches"] > 0:
        kname = ("fdtd::stepK_lean_split<%d>: split field storage (fp32 hi + bf16 lo planes, 6 B per field value, float64 "
                 "coefficients and arithmetic): 52 algorithmic bytes per cell-update; one launch advances its tiles %d steps"
                 % (args.temporal, args.temporal))
    achieved = alg_bytes / (launch_ms * 1e-3) / 1e9
    line = {
        "metric": METRIC, "value": value, "unit": "Gcell-updates/s", "n_gpus": n_gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True,
        "scaling": "strong" if args.total_rows else "weak",
        "vs_baseline": None, "dtype": args.dtype, "data": "synthetic", "config": workload_config(args, n_gpus, e),
        "clocks": clocks, "e2e": e2e, "gpu_launches": int(launches),
        "repeats": {"n": reps, "ms_per_step": [m / args.steps for m in rep_ms],
                    "value_min": cells * args.steps / (max(rep_ms) * 1e-3) / 1e9,
                    "value_max": cells * args.steps / (min(rep_ms) * 1e-3) / 1e9,
                    "what": "value = median of n back-to-back repetitions of the timed --steps region",
                    "warmup_steps_run": args.warmup + structural},
        "roofline": {"bound": "hbm", "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                     "traffic": None, "peak_source": peak_src,
                     "algorithmic_bytes_per_launch": alg_bytes,
                     "kernel": kname,
                     "launch_ms": launch_ms,
                     "launch_ms_min_max": [kt["min_ms"], kt["max_ms"]] if kt["launches"] else None,
                     "launches_timed": kt["launches"],
                     "kernel_share_of_step": (kt["total_ms"] / (sum(rep_ms)) if kt["launches"] else 1.0),
                     "steps_per_launch": args.temporal,
                     "whole_step_achieved": value / n_gpus * bpc, "whole_step_frac": value / n_gpus * bpc / peak},
        "wall_s_timed_region": wall,
    }
    if with_snapshots is not None:
        line["with_snapshots"] = with_snapshots
    if n_gpus > 1:
        line["halo"] = {"bytes_sent_per_step_per_gpu": halo_bytes / max(1, st["steps"]),
                        "nvlink_frac_of_900GBs": (halo_bytes / max(1, st["steps"])) / (ms / args.steps * 1e-3) / 900e9}
    traffic_file = os.path.join(ROOT, "profiles", "traffic_%s%s%s.json" % (
        args.dtype, "_temporal%d" % args.temporal if args.temporal > 1 else "",
        (("_skew_umu" if UMU_KERNEL == "stepK_skew_umu" else "_lean_umu") if e.uniform_mu() else "_lean") if e.kstep_variant == 1 else ""))
    if os.path.isfile(traffic_file):
        try:
            rf = line["roofline"]
            tf = json.load(open(traffic_file))
            rf["traffic"] = tf["dram_bytes_per_launch"]
            if tf.get("algorithmic_bytes_per_launch") and alg_bytes < 0.98 * tf["algorithmic_bytes_per_launch"]:
                # the captured launch covered the whole slab; this run's launches cover fewer tiles each (deep halos on
                # y-slabs: interior tile rows and interface tile rows are separate launches) -- same bytes per tile
                rf["traffic"] = tf["dram_bytes_per_launch"] * alg_bytes / tf["algorithmic_bytes_per_launch"]
                rf["traffic_scaled_from"] = tf["dram_bytes_per_launch"]
            # DRAM-level view of the same launch: the ncu bytes over the live launch time.  With temporal blocking the
            # algorithmic fraction above exceeds 1 by design; THIS one cannot, and says how close the pass is to the bus.
            rf["dram_achieved"] = rf["traffic"] / (rf["launch_ms"] * 1e-3) / 1e9          # the kernel's own event time
            rf["dram_frac"] = rf["dram_achieved"] / peak
        except Exception:
            pass
    if n_gpus == 1 and not args.no_cpu:
        v, n, dt, kind = cpu_baseline()
        line["cpu_baseline"] = {"value": v, "unit": "Gcell-updates/s", "cores": 1, "kind": kind,
                                "sample": "%d x %d slab of the C5 recipe, %d numpy steps in %.1f s (host has %d cores; "
                                          "numpy ufuncs use 1)" % (CPU_SAMPLE + (n, dt, os.cpu_count())),
                                "threaded_c_port": cpu_baseline_c()}
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
