#!/bin/bash
# Round 2, call E (1 GPU): stepK_edge per tile class (edge.cuh EDGE_CLASSES) and the library split into translation units.
# GPU suite, the class instantiations against the all-features one on the C5 slab, launch list, compute-sanitizer logs.
mkdir -p gpurun_out
O=gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > ${O}_gpu_tests.log 2>&1; echo "rc=$?" >> ${O}_gpu_tests.log
tail -4 ${O}_gpu_tests.log
timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_classes.json 2> ${O}_bench_classes.err; echo "classes rc=$?"
FDTD_EDGE_GENERIC=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_generic.json 2> ${O}_bench_generic.err; echo "generic rc=$?"
timeout 400 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "default rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${O}_launches.csv python bench.py --no-cpu --no-e2e --steps 16 --warmup 3 > ${O}_ncu.log 2>&1; echo "ncu rc=$?"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 python tools/sanitizer_scene.py > ${O}_sanitizer_memcheck.log 2>&1; echo "memcheck rc=$?"
timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 python tools/sanitizer_scene.py > ${O}_sanitizer_racecheck.log 2>&1; echo "racecheck rc=$?"
timeout 600 compute-sanitizer --tool synccheck --error-exitcode 9 python tools/sanitizer_scene.py stepK_stream stepK_lean_edge step_resident > ${O}_sanitizer_synccheck.log 2>&1; echo "synccheck rc=$?"
tail -3 ${O}_sanitizer_*.log
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2e_bench_*.json")):
    try:
        d = json.load(open(f)); r = d.get("roofline") or {}
        print(f.split("r2e_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "share", r.get("kernel_share_of_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
