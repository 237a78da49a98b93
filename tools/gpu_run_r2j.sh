#!/bin/bash
# Round 2, call J (1 GPU): split field storage with its plain K-step kernel (stepK_lean_split for the plain tiles,
# stepK_edge<.., split> for the others): parity tests, bench lines (default geometry + tile heights), launch list and one
# ncu --set full capture of the new kernel; the driver's own 20-step command lines for both arms.
mkdir -p gpurun_out
O=gpurun_out/r2j
timeout 900 python -m pytest tests/test_gpu_split_storage.py -q -s -p no:cacheprovider --durations=6 > ${O}_split_storage_tests.log 2>&1; echo "rc=$?" >> ${O}_split_storage_tests.log
tail -8 ${O}_split_storage_tests.log
timeout 300 python bench.py --dtype f32x --no-cpu --steps 400 > ${O}_bench_f32x.json 2> ${O}_bench_f32x.err; echo "f32x rc=$?"
for tr in 64 128 256; do
  timeout 300 python bench.py --dtype f32x --no-cpu --no-e2e --steps 200 --tile-rows $tr --repeats 3 > ${O}_bench_f32x_tr$tr.json 2> ${O}_bench_f32x_tr$tr.err; echo "f32x tr$tr rc=$?"
done
timeout 300 python bench.py --dtype f32x --no-cpu --no-e2e --steps 200 --temporal 6 --repeats 3 > ${O}_bench_f32x_k6.json 2> ${O}_bench_f32x_k6.err; echo "f32x k6 rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver_cmd_20steps.json 2> ${O}_bench_driver_cmd_20steps.err; echo "driver cmd rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver_cmd_reference.json 2> ${O}_bench_driver_cmd_reference.err; echo "driver ref rc=$?"
CMD="python bench.py --dtype f32x --no-cpu --no-e2e --steps 16 --warmup 8 --repeats 1"
timeout 300 $CMD > ${O}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file ${O}_launches_f32x.csv $CMD > ${O}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stepK_lean_split -s 3 -c 1 -o ${O}_prof_lean_split $CMD > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i ${O}_prof_lean_split.ncu-rep --page raw --csv > ${O}_ncu_full_stepK_lean_split_raw.csv 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2j_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2j_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "frac", r.get("frac"), "share", r.get("kernel_share_of_step"))
    except Exception as e:
        print(f, "FAILED", e)
PY
