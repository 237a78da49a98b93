#!/bin/bash
# Round 2, call L (1 GPU): split storage, the bf16 half of the per-stage rounding on the FP64 pipe for some of a stage's six
# values (FDTD_SPLIT_XU_PER_STAGE = how many keep the two XU conversions): parity tests with the default build (4), then the
# sweep over side-by-side libraries built with 0 / 2 / 3 / 5 / 6.
mkdir -p gpurun_out
O=gpurun_out/r2l
timeout 900 python -m pytest tests/test_gpu_split_storage.py -q -s -p no:cacheprovider --durations=4 > ${O}_split_storage_tests.log 2>&1; echo "rc=$?" >> ${O}_split_storage_tests.log
tail -6 ${O}_split_storage_tests.log
timeout 300 python bench.py --dtype f32x --no-cpu --no-e2e --steps 200 --repeats 3 > ${O}_bench_f32x_xu4.json 2> ${O}_bench_f32x_xu4.err; echo "xu4 rc=$?"
for x in 0 2 3 5 6; do
  FDTD_LIB_SUFFIX=_xu$x FDTD_NVCC_EXTRA="-DFDTD_SPLIT_XU_PER_STAGE=$x" timeout 300 python bench.py --dtype f32x --no-cpu --no-e2e --steps 200 --repeats 3 > ${O}_bench_f32x_xu$x.json 2> ${O}_bench_f32x_xu$x.err; echo "xu$x rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2l_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("r2l_bench_")[1], round(d["value"], 2), "kernel_ms", r.get("launch_ms"), "frac", r.get("frac"), d["clocks"])
    except Exception as e:
        print(f, "FAILED", e)
PY
