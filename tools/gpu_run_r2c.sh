#!/bin/bash
# Round 2, third single-GPU call: tall plain tiles with short edge tiles; proper ncu captures of the K = 8 kernels.
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_frame_window.py tests/test_gpu_deep_halo.py tests/test_gpu_kstep_lean.py -q -p no:cacheprovider -x \
    > gpurun_out/r2c_tests.log 2>&1; echo "rc=$?" >> gpurun_out/r2c_tests.log
tail -2 gpurun_out/r2c_tests.log
B="python bench.py --no-cpu --no-e2e --steps 120 --warmup 6 --repeats 3 --frame-mode 3 --kstep-variant 1"
run() { name=$1; shift; timeout 120 $B "$@" > gpurun_out/r2c_$name.json 2> gpurun_out/r2c_$name.err; python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2c_$name.json")); r = d["roofline"]
    print("$name", round(d["value"], 1), "Gcell/s  kernel %.3f ms x%d  share %.2f  launches %d  range %.1f-%.1f" % (
        r["launch_ms"], r["launches_timed"], r["kernel_share_of_step"], d["gpu_launches"], d["repeats"]["value_min"], d["repeats"]["value_max"]))
except Exception as e:
    print("$name FAILED", e)
PY
}
for tr in 128 256 512 1024 2048 8192; do run k8_tr$tr --temporal 8 --tile-rows $tr; done
for tr in 256 1024; do run k7_tr$tr --temporal 7 --tile-rows $tr; done
for tr in 256 1024; do run k6_tr$tr --temporal 6 --tile-rows $tr; done
for er in 16 64; do FDTD_EDGE_ROWS=$er run k8_tr512_er$er --temporal 8 --tile-rows 512; done
run k8_tr512_w2 --temporal 8 --tile-rows 512 --warps 2
run k8_tr512_w4 --temporal 8 --tile-rows 512 --warps 4
run f32_k8_tr512 --temporal 8 --tile-rows 512 --dtype f32
run f32_k8_tr128 --temporal 8 --tile-rows 128 --dtype f32
run k8_tr512_1000steps --temporal 8 --tile-rows 512 --steps 1000
# ncu (after the plain runs): launch list of our kernels, then full captures of the K = 8 instances (warm-up of 8 steps
# = one K = 8 group, so the first matching launch is the K = 8 kernel)
N="python bench.py --no-cpu --no-e2e --steps 24 --warmup 8 --repeats 1 --temporal 8 --tile-rows 512 --frame-mode 3 --kstep-variant 1"
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'step|plane|synth' -c 40 --csv \
    --log-file gpurun_out/r2c_launches_k8_lean_edge.csv $N > gpurun_out/r2c_ncu_launches.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stepK_lean -c 1 -o gpurun_out/r2c_stepK_lean_k8 $N > gpurun_out/r2c_ncu_lean.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:stepK_edge -c 1 -o gpurun_out/r2c_stepK_edge_k8 $N > gpurun_out/r2c_ncu_edge.log 2>&1
