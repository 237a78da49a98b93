#!/bin/bash
# Round 2, call M (1 GPU): stepK_lean_umu (the plain K-step kernel without the mu_coef stream when that plane is uniform):
# the whole GPU suite with this build (StepParams grew, so every kernel's parameter offsets moved), A/B bench lines,
# launch list and ncu --set full of the new kernel.
mkdir -p gpurun_out
O=gpurun_out/r2m
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=8 > ${O}_gpu_tests.log 2>&1; echo "rc=$?" >> ${O}_gpu_tests.log
tail -4 ${O}_gpu_tests.log
timeout 400 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "default rc=$?"
FDTD_UNIFORM_MU=0 timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_general_kernel.json 2> ${O}_bench_general_kernel.err; echo "general rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_umu.json 2> ${O}_bench_umu.err; echo "umu rc=$?"
timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 > ${O}_bench_driver_cmd_20steps.json 2> ${O}_bench_driver_cmd_20steps.err; echo "driver cmd rc=$?"
timeout 300 python bench.py --dtype f32 --no-cpu --no-e2e --steps 400 > ${O}_bench_f32_umu.json 2> ${O}_bench_f32_umu.err; echo "f32 rc=$?"
FDTD_UNIFORM_MU=0 timeout 300 python bench.py --dtype f32 --no-cpu --no-e2e --steps 400 > ${O}_bench_f32_general.json 2> ${O}_bench_f32_general.err; echo "f32 general rc=$?"
timeout 300 python bench.py --dtype f32x --no-cpu --steps 400 > ${O}_bench_f32x.json 2> ${O}_bench_f32x.err; echo "f32x rc=$?"
CMD="python bench.py --no-cpu --no-e2e --steps 16 --warmup 8 --repeats 1"
timeout 300 $CMD > ${O}_plain.log 2>&1 &&
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file ${O}_launches_f64.csv $CMD > ${O}_ncu_list.log 2>&1; echo "ncu list rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:stepK_lean_umu -s 3 -c 1 -o ${O}_prof_lean_umu $CMD > ${O}_ncu_full.log 2>&1; echo "ncu full rc=$?"
ncu -i ${O}_prof_lean_umu.ncu-rep --page raw --csv > ${O}_ncu_full_stepK_lean_umu_raw.csv 2>/dev/null
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2m_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2m_bench_")[1], round(d["value"], 2), "e2e", (d.get("e2e") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "frac", r.get("frac"), "umu", d["config"].get("uniform_mu"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "FAILED", e)
PY
