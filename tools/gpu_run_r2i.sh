#!/bin/bash
# Round 2, call I (1 GPU): first hardware run of split field storage (dtype "float32x": fp32 hi + bf16 lo planes,
# float64 arithmetic) -- its GPU tests incl. the 1e-5 gate on every snapshot of C1/C3/C4 and 12 000 calls of C2, then the
# whole GPU suite with the build that carries it, the default bench line, the f32x and f32 bench lines and the launch list.
mkdir -p gpurun_out
O=gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_split_storage.py -q -s -p no:cacheprovider --durations=6 > ${O}_split_storage_tests.log 2>&1; echo "rc=$?" >> ${O}_split_storage_tests.log
tail -12 ${O}_split_storage_tests.log
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 --deselect tests/test_gpu_split_storage.py > ${O}_gpu_tests.log 2>&1; echo "rc=$?" >> ${O}_gpu_tests.log
tail -4 ${O}_gpu_tests.log
timeout 400 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --dtype f32x --no-cpu --steps 200 > ${O}_bench_f32x.json 2> ${O}_bench_f32x.err; echo "f32x rc=$?"
timeout 300 python bench.py --dtype f32 --no-cpu --steps 400 > ${O}_bench_f32.json 2> ${O}_bench_f32.err; echo "f32 rc=$?"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file ${O}_launches_f32x.csv python bench.py --dtype f32x --no-cpu --no-e2e --steps 16 --warmup 8 > ${O}_ncu_f32x.log 2>&1; echo "ncu rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2i_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2i_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "frac", r.get("frac"), "share", r.get("kernel_share_of_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
