#!/bin/bash
# Round 2, call P (1 GPU): stepK_skew_umu (the stage chain of the K-step pipeline cut at skewed boundaries): parity tests,
# then A/B bench lines: lean vs skew (default mask 0x48), other masks and row-loop unroll 4 as side-by-side libraries.
mkdir -p gpurun_out
O=gpurun_out/r2p
timeout 900 python -m pytest tests/test_gpu_uniform_mu.py -q -p no:cacheprovider --durations=3 > ${O}_uniform_mu_tests.log 2>&1; echo "rc=$?" >> ${O}_uniform_mu_tests.log
tail -4 ${O}_uniform_mu_tests.log
B="python bench.py --no-cpu --no-e2e --steps 400 --repeats 3"
timeout 300 $B > ${O}_bench_lean_umu.json 2> ${O}_bench_lean_umu.err; echo "lean rc=$?"
FDTD_SKEW=1 timeout 300 $B > ${O}_bench_skew48.json 2> ${O}_bench_skew48.err; echo "skew48 rc=$?"
for v in sk10 sk54 sk7e sk28; do
  m=$(echo $v | sed 's/sk/0x/')
  FDTD_SKEW=1 FDTD_LIB_SUFFIX=_$v FDTD_NVCC_EXTRA="-DFDTD_SKEW_MASK=$(echo $m | tr a-f A-F | sed 's/0X/0x/')" timeout 300 $B > ${O}_bench_$v.json 2> ${O}_bench_$v.err; echo "$v rc=$?"
done
FDTD_LIB_SUFFIX=_u4 FDTD_NVCC_EXTRA="-DFDTD_LEAN_UNROLL=4" timeout 300 $B > ${O}_bench_lean_umu_u4.json 2> ${O}_bench_lean_umu_u4.err; echo "u4 lean rc=$?"
FDTD_SKEW=1 FDTD_LIB_SUFFIX=_u4 FDTD_NVCC_EXTRA="-DFDTD_LEAN_UNROLL=4" timeout 300 $B > ${O}_bench_skew48_u4.json 2> ${O}_bench_skew48_u4.err; echo "u4 skew rc=$?"
FDTD_UNIFORM_MU=0 FDTD_LIB_SUFFIX=_u4 FDTD_NVCC_EXTRA="-DFDTD_LEAN_UNROLL=4" timeout 300 $B > ${O}_bench_lean_general_u4.json 2> ${O}_bench_lean_general_u4.err; echo "u4 general rc=$?"
FDTD_SKEW=1 timeout 300 $B --dtype f32 > ${O}_bench_f32_skew48.json 2> ${O}_bench_f32_skew48.err; echo "f32 skew rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2p_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2p_bench_")[1], round(d["value"], 2), "kernel_ms", r.get("launch_ms"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "FAILED", e)
PY
