#!/bin/bash
# Round 2, last multi-GPU call (8 GPUs, charged 8x, so only the essentials and the most important first):
# weak scaling at 1000 steps, the same box's 1-GPU line, BASELINE config 5 as written (65536^2, 10 000 steps), the driver's
# 20-step command, then 1 GPU == 8 GPUs bitwise (tools/slab_check.py --big) -- all with the final build (stepK_skew_umu).
N=${1:-8}
mkdir -p gpurun_out
O=gpurun_out/r2r_n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 200 $TR bench.py --gpus $N --no-cpu --no-e2e > ${O}_bench_weak_f64.json 2> ${O}_bench_weak_f64.err; echo "weak f64 rc=$?"
timeout 120 python bench.py --no-cpu --no-e2e --steps 400 --repeats 3 > ${O}_bench_samebox_n1_f64.json 2> ${O}_bench_samebox_n1_f64.err; echo "n1 rc=$?"
timeout 200 $TR bench.py --gpus $N --no-cpu --no-e2e --steps 10000 --repeats 1 > ${O}_bench_weak_f64_10000steps.json 2> ${O}_bench_weak_f64_10000steps.err; echo "10000 steps rc=$?"
timeout 120 $TR bench.py --gpus $N --no-cpu --no-e2e --steps 20 --warmup 5 > ${O}_bench_weak_f64_20steps.json 2> ${O}_bench_weak_f64_20steps.err; echo "20 steps rc=$?"
python - <<PY
import json, glob
for f in sorted(glob.glob("${O}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
        print(f.split("_bench_")[1], "n", d["n_gpus"], round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "umu", d["config"].get("uniform_mu"), d["repeats"]["value_min"], d["repeats"]["value_max"])
    except Exception as e:
        print(f, "FAILED", e)
PY
timeout 400 $TR tools/slab_check.py --big > ${O}_slab_check.log 2>&1; echo "slab_check rc=$?"; tail -2 ${O}_slab_check.log | cut -c1-200
