#!/bin/bash
# Round 2, call K (1 GPU): split storage with the conversion-free rounding (split_round_hi: the fp32 rounding of a stage's
# outputs on the FP64 pipe instead of the XU pipe) -- parity tests, bench line, launch list, ncu --set full of
# stepK_lean_split; and the float64 run with snapshots now that the host-side snapshot copy is multi-threaded.
mkdir -p gpurun_out
O=gpurun_out/r2k
timeout 900 python -m pytest tests/test_gpu_split_storage.py -q -s -p no:cacheprovider --durations=4 > ${O}_split_storage_tests.log 2>&1; echo "rc=$?" >> ${O}_split_storage_tests.log
tail -8 ${O}_split_storage_tests.log
timeout 300 python bench.py --dtype f32x --no-cpu --steps 400 > ${O}_bench_f32x.json 2> ${O}_bench_f32x.err; echo "f32x rc=$?"
timeout 300 python bench.py --dtype f32x --no-cpu --no-e2e --steps 200 --tile-rows 256 --repeats 3 > ${O}_bench_f32x_tr256.json 2> ${O}_bench_f32x_tr256.err; echo "f32x tr256 rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 240 --snapshot-every 120 --repeats 3 > ${O}_bench_f64_snap120.json 2> ${O}_bench_f64_snap120.err; echo "snap120 rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 240 --snapshot-every 60 --repeats 3 > ${O}_bench_f64_snap60.json 2> ${O}_bench_f64_snap60.err; echo "snap60 rc=$?"
FDTD_SNAPSHOT_THREADS=1 timeout 300 python bench.py --no-cpu --no-e2e --steps 240 --snapshot-every 60 --repeats 3 > ${O}_bench_f64_snap60_1thread.json 2> ${O}_bench_f64_snap60_1thread.err; echo "snap60 1 thread rc=$?"
timeout 300 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider -k "scenarios_through_the_python_api or snapshot" > ${O}_snapshot_parity_tests.log 2>&1; echo "rc=$?" >> ${O}_snapshot_parity_tests.log; tail -2 ${O}_snapshot_parity_tests.log
CMD="python bench.py --dtype f32x --no-cpu --no-e2e --steps 16 --warmup 8 --repeats 1"
timeout 300 $CMD > ${O}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file ${O}_launches_f32x.csv $CMD > ${O}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stepK_lean_split -s 3 -c 1 -o ${O}_prof_lean_split $CMD > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i ${O}_prof_lean_split.ncu-rep --page raw --csv > ${O}_ncu_full_stepK_lean_split_raw.csv 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2k_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2k_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "frac", r.get("frac"), "snap", (d.get("with_snapshots") or {}).get("value"))
    except Exception as e:
        print(f, "FAILED", e)
PY
