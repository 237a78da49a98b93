#!/bin/bash
# Round 2, call D (1 GPU): the whole GPU suite with the new defaults (K = 8, stepK_lean + stepK_edge, scratch-free), then
# the default bench line, the driver-style 20-step line, the reference arm and the snapshot legs.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=12 > gpurun_out/r2d_gpu_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2d_gpu_tests.log
tail -4 gpurun_out/r2d_gpu_tests.log
timeout 400 python bench.py > gpurun_out/r2d_bench_default.json 2> gpurun_out/r2d_bench_default.err; echo "default rc=$?"
timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu > gpurun_out/r2d_bench_20steps.json 2> gpurun_out/r2d_bench_20steps.err; echo "20steps rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2d_bench_reference.json 2> gpurun_out/r2d_bench_reference.err; echo "ref rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 240 --snapshot-every 120 > gpurun_out/r2d_bench_snap120.json 2> gpurun_out/r2d_bench_snap120.err; echo "snap120 rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 240 --snapshot-every 60 > gpurun_out/r2d_bench_snap60.json 2> gpurun_out/r2d_bench_snap60.err; echo "snap60 rc=$?"
timeout 200 python bench.py --no-cpu --no-e2e --dtype f32 > gpurun_out/r2d_bench_f32.json 2> gpurun_out/r2d_bench_f32.err; echo "f32 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_bench_*.json")):
    try:
        d = json.load(open(f)); r = d.get("roofline") or {}
        print(f.split("r2d_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "snap", (d.get("with_snapshots") or {}).get("value"), "cpu", (d.get("cpu_baseline") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
