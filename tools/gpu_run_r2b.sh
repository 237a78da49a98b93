#!/bin/bash
# Round 2, second single-GPU call: first run of stepK_edge (frame mode 3) -- parity, then what it buys.
mkdir -p gpurun_out
timeout 420 python -m pytest tests/test_gpu_frame_window.py tests/test_gpu_deep_halo.py -q -p no:cacheprovider -x --durations=5 \
    > gpurun_out/r2b_edge_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2b_edge_tests.log
tail -3 gpurun_out/r2b_edge_tests.log
B="python bench.py --no-cpu --no-e2e --steps 120 --warmup 6 --repeats 3"
run() { name=$1; shift; timeout 120 $B "$@" > gpurun_out/r2b_$name.json 2> gpurun_out/r2b_$name.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2b_$name.json")); r = d["roofline"]
    print("$name", round(d["value"], 1), "Gcell/s  kernel %.3f ms x%d  share %.2f  launches %d  range %.1f-%.1f" % (
        r["launch_ms"], r["launches_timed"], r["kernel_share_of_step"], d["gpu_launches"], d["repeats"]["value_min"], d["repeats"]["value_max"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
run k6_stream_mode1 --temporal 6 --frame-mode 1
run k6_stream_mode3 --temporal 6 --frame-mode 3
run k6_lean_mode3 --temporal 6 --frame-mode 3 --kstep-variant 1
run k7_lean_mode3 --temporal 7 --frame-mode 3 --kstep-variant 1
run k8_lean_mode3 --temporal 8 --frame-mode 3 --kstep-variant 1
run k8_lean_mode1 --temporal 8 --frame-mode 1 --kstep-variant 1
run k8_stream_mode3 --temporal 8 --frame-mode 3
for tr in 32 96 128 256 1024; do run k8_lean_mode3_tr$tr --temporal 8 --frame-mode 3 --kstep-variant 1 --tile-rows $tr; done
for tr in 96 256; do run k6_lean_mode3_tr$tr --temporal 6 --frame-mode 3 --kstep-variant 1 --tile-rows $tr; done
run k8_lean_mode3_w2 --temporal 8 --frame-mode 3 --kstep-variant 1 --warps 2
run f32_k8_lean_mode3 --temporal 8 --frame-mode 3 --kstep-variant 1 --dtype f32
run f32_k6_mode3 --temporal 6 --frame-mode 3 --dtype f32
# launch list + full captures (only after the plain runs above)
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2b_launches_k8_lean_mode3.csv \
    python bench.py --no-cpu --no-e2e --steps 24 --warmup 3 --repeats 1 --temporal 8 --frame-mode 3 --kstep-variant 1 > gpurun_out/r2b_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stepK_lean -c 1 -o gpurun_out/r2b_stepK_lean_k8 \
    python bench.py --no-cpu --no-e2e --steps 16 --warmup 3 --repeats 1 --temporal 8 --frame-mode 3 --kstep-variant 1 > gpurun_out/r2b_ncu_lean.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stepK_edge -c 1 -o gpurun_out/r2b_stepK_edge_k8 \
    python bench.py --no-cpu --no-e2e --steps 16 --warmup 3 --repeats 1 --temporal 8 --frame-mode 3 --kstep-variant 1 > gpurun_out/r2b_ncu_edge.log 2>&1
