#!/bin/bash
# Round 2, call H (1 GPU): fdtd_run_streamed -- parity tests, then the e2e leg of bench.py at 20 / 200 / 1000 steps for
# a few band counts and wing lengths (FDTD_STREAM_BANDS / FDTD_STREAM_WING).
mkdir -p gpurun_out
O=gpurun_out/r2h
timeout 300 python -m pytest tests/test_gpu_streamed.py tests/test_gpu_realtime.py -q -p no:cacheprovider > ${O}_streamed_tests.log 2>&1; echo rc=$? >> ${O}_streamed_tests.log; tail -5 ${O}_streamed_tests.log
for cfg in "20 8 7" "20 16 15" "20 4 3" "200 8 7" "200 16 15" "200 8 3" "1000 8 7" "1000 8 3" "1000 16 4"; do
  set -- $cfg
  FDTD_STREAM_BANDS=$2 FDTD_STREAM_WING=$3 timeout 300 python bench.py --no-cpu --steps $1 --warmup 5 --repeats 2 > ${O}_bench_s$1_b$2_w$3.json 2> ${O}_bench_s$1_b$2_w$3.err; echo "steps $1 bands $2 wing $3 rc=$?"
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2h_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); e = d["e2e"]
        print(f.split("r2h_bench_")[1], "device", round(d["value"], 1), "e2e", round(e["value"], 1), round(e["seconds"], 4), "sequential", round(e["sequential"]["value"], 1), round(e["sequential"]["seconds"], 4))
    except Exception as ex:
        print(f, "FAILED", ex)
PY
