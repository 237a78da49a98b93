#!/bin/bash
# Round 2, call Q (1 GPU): the final build -- stepK_skew_umu default for a uniform mu plane, row loop of the K >= 7 lean /
# skew pipelines unrolled by 4 (every such kernel's SASS changed): whole GPU suite, smoke(), the bench lines, launch list and
# ncu --set full of the dominant kernel.
mkdir -p gpurun_out
O=gpurun_out/r2q
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > ${O}_gpu_tests.log 2>&1; echo "rc=$?" >> ${O}_gpu_tests.log
tail -3 ${O}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log | cut -c1-200
timeout 400 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver_cmd_20steps.json 2> ${O}_bench_driver_cmd_20steps.err; echo "driver cmd rc=$?"
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver_cmd_reference.json 2> ${O}_bench_driver_cmd_reference.err; echo "driver ref rc=$?"
FDTD_UNIFORM_MU=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_general_kernel.json 2> ${O}_bench_general_kernel.err; echo "general rc=$?"
timeout 300 python bench.py --dtype f32 --no-cpu --steps 400 > ${O}_bench_f32.json 2> ${O}_bench_f32.err; echo "f32 rc=$?"
timeout 300 python bench.py --dtype f32x --no-cpu --steps 400 > ${O}_bench_f32x.json 2> ${O}_bench_f32x.err; echo "f32x rc=$?"
CMD="python bench.py --no-cpu --no-e2e --steps 16 --warmup 8 --repeats 1"
timeout 300 $CMD > ${O}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file ${O}_launches_f64.csv $CMD > ${O}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stepK_skew_umu -s 3 -c 1 -o ${O}_prof_skew_umu $CMD > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i ${O}_prof_skew_umu.ncu-rep --page raw --csv > ${O}_ncu_full_stepK_skew_umu_raw.csv 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2q_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}; e = d.get("e2e") or {}
        print(f.split("r2q_bench_")[1], round(d["value"], 2), "e2e", e.get("value"), "kernel_ms", r.get("launch_ms"), "frac", r.get("frac"),
              (d.get("clocks") or {}).get("sm_mhz"), (d.get("clocks") or {}).get("reasons"))
    except Exception as e:
        print(f, "FAILED", e)
PY
