#!/bin/bash
# Round 2, final multi-GPU call:  gpurun --gpus N -- 'bash tools/gpu_run_multi_r2o.sh N'
# The final build (stepK_lean_umu for the C5 recipe's uniform mu plane) at N GPUs: 1 GPU == N GPUs bitwise
# (tools/slab_check.py --big), the multi-GPU pytest cases (N <= 4), an N = 1 line from the same box, weak scaling at 1000
# steps and with the driver's 20-step command, and at N = 8 BASELINE config 5 as written (65536^2, 10 000 steps).
N=${1:-2}
mkdir -p gpurun_out
O=gpurun_out/r2o_n${N}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531"
timeout 900 $TR tools/slab_check.py --big > ${O}_slab_check.log 2>&1; echo "slab_check rc=$?"; tail -3 ${O}_slab_check.log | cut -c1-200
if [ "$N" -le 4 ]; then
timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -k "nccl or sharded or torchrun" > ${O}_multi_gpu_tests.log 2>&1; echo "multi tests rc=$?"; tail -2 ${O}_multi_gpu_tests.log
fi
timeout 300 python bench.py --no-cpu --no-e2e --steps 400 > ${O}_bench_samebox_n1_f64.json 2> ${O}_bench_samebox_n1_f64.err; echo "n1 rc=$?"
timeout 400 $TR bench.py --gpus $N --no-cpu > ${O}_bench_weak_f64.json 2> ${O}_bench_weak_f64.err; echo "weak f64 rc=$?"
timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --steps 20 --warmup 5 > ${O}_bench_weak_f64_20steps.json 2> ${O}_bench_weak_f64_20steps.err; echo "20 steps rc=$?"
timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --dtype f32 > ${O}_bench_weak_f32.json 2> ${O}_bench_weak_f32.err; echo "weak f32 rc=$?"
if [ "$N" -ge 8 ]; then
  timeout 400 $TR bench.py --gpus $N --no-cpu --no-e2e --steps 10000 --repeats 1 > ${O}_bench_weak_f64_10000steps.json 2> ${O}_bench_weak_f64_10000steps.err; echo "10000 steps rc=$?"
  timeout 300 $TR bench.py --gpus $N --no-cpu --no-e2e --total-rows 65536 --steps 200 > ${O}_bench_strong_f64.json 2> ${O}_bench_strong_f64.err; echo "strong f64 rc=$?"
fi
python - <<PY
import json, glob
for f in sorted(glob.glob("${O}_bench_*.json")):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1]); r = d.get("roofline") or {}
        print(f.split("_bench_")[1], "n", d["n_gpus"], round(d["value"], 1), "ms/step", round(d["ms_per_step"], 4), "e2e", (d.get("e2e") or {}).get("value"),
              "umu", d["config"].get("uniform_mu"), "dram_frac", r.get("dram_frac"), "halo", (d.get("halo") or {}))
    except Exception as e:
        print(f, "FAILED", e)
PY
