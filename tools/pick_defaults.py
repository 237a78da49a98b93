"""After `tools/next_gpu_run.sh` / `next_gpu_run_multi.sh`: one table of every bench line they left in gpurun_out/, the
parity legs' verdicts, and which opt-in kernels earned a default flip (an opt-in replaces the validated default only if
its parity leg passed AND it is faster by more than run-to-run noise, 2 %).

    python tools/pick_defaults.py [gpurun_out]
"""
import glob
import json
import os
import sys

d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "gpurun_out")
rows = []
for path in sorted(glob.glob(os.path.join(d, "next_bench_*.json")) + glob.glob(os.path.join(d, "multi_*_n*.json"))):
    try:
        line = json.loads(open(path).read().strip().splitlines()[-1])
    except Exception as exc:                                   # an empty file = the leg failed; its .err says why
        rows.append((os.path.basename(path), None, "no JSON line (%s)" % type(exc).__name__))
        continue
    c = line.get("config", {})
    rows.append((os.path.basename(path), line["value"],
                 "N=%d %s K=%s frame_mode=%s kstep_variant=%s halo_depth=%s %.3f ms/step" % (
                     line["n_gpus"], line["dtype"], c.get("temporal_blocking"), c.get("frame_mode"),
                     c.get("kstep_variant"), c.get("halo_depth"), line["ms_per_step"])))
for name, value, what in rows:
    print("%-44s %10s  %s" % (name, "-" if value is None else "%.1f" % value, what))

for log in ("next_new_kernels_tests.log", "multi_slab_check.log", "multi_solver_sharded.log"):
    path = os.path.join(d, log)
    if os.path.exists(path):
        tail = open(path).read().strip().splitlines()[-3:]
        print("\n%s:\n  %s" % (log, "\n  ".join(tail)))

by = {name: value for name, value, _ in rows if value is not None}
base = by.get("next_bench_frame_mode0.json")
if base:
    print("\nbaseline (validated defaults, K = 6): %.1f Gcell/s" % base)
    for name, flip in (("next_bench_frame_mode1.json", "engine.py: FDTD_FRAME_MODE default 1"),
                       ("next_bench_frame_mode2.json", "engine.py: FDTD_FRAME_MODE default 2"),
                       ("next_bench_kstepvariant1.json", "engine.py: FDTD_KSTEP_VARIANT default 1"),
                       ("next_bench_kstepvariant1framemode1.json", "both of the above"),
                       ("next_bench_kstepvariant1temporal8.json", "bench.py / fdtd_set_temporal_blocking: K = 8 with stepK_lean")):
        if name in by:
            gain = by[name] / base - 1
            print("  %-44s %+6.1f %%  %s" % (name, 100 * gain, ("FLIP: " + flip) if gain > 0.02 else "keep the default"))
    print("(flip only with the matching parity leg green; then `python tools/sass_manifest.py --write <kernel>`)")
