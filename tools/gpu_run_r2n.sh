#!/bin/bash
# Round 2, call N (1 GPU): the whole GPU suite and smoke() with the final host side (streamed runs scan the mu plane behind
# their first groups), the default bench line incl. e2e, and an occupancy sweep of stepK_lean_umu<.., 8>: registers capped
# at 200 / 184 / 168 (10 / 11 / 12 warps per SM instead of 9) as side-by-side libraries.
mkdir -p gpurun_out
O=gpurun_out/r2n
timeout 1500 python -m pytest tests -m gpu -q -p no:cacheprovider --durations=5 > ${O}_gpu_tests.log 2>&1; echo "rc=$?" >> ${O}_gpu_tests.log
tail -3 ${O}_gpu_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > ${O}_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 ${O}_smoke.log | cut -c1-200
timeout 400 python bench.py > ${O}_bench_default.json 2> ${O}_bench_default.err; echo "default rc=$?"
timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_r217.json 2> ${O}_bench_r217.err; echo "r217 rc=$?"
for n in 200 184 168; do
  FDTD_LIB_SUFFIX=_r$n FDTD_NVCC_EXTRA="-DFDTD_UMU_MAXNREG=$n" timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_r$n.json 2> ${O}_bench_r$n.err; echo "r$n rc=$?"
done
FDTD_LIB_SUFFIX=_r184 FDTD_NVCC_EXTRA="-DFDTD_UMU_MAXNREG=184" timeout 300 python bench.py --no-cpu --no-e2e --steps 400 --tile-rows 256 > ${O}_bench_r184_tr256.json 2> ${O}_bench_r184_tr256.err; echo "r184 tr256 rc=$?"
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}; e = d.get("e2e") or {}
        print(f.split("r2n_bench_")[1], round(d["value"], 2), "e2e", e.get("value"), "seq", (e.get("sequential") or {}).get("value"), "kernel_ms", r.get("launch_ms"),
              "umu", d["config"].get("uniform_mu"), d["clocks"]["sm_mhz"], d["clocks"]["reasons"])
    except Exception as e:
        print(f, "FAILED", e)
PY
