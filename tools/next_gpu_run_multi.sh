#!/bin/bash
# Multi-GPU measurements still missing from SURVEY.md 8(d) after round 1 (charged N x box time: run once, late in a round):
#   gpurun --gpus 8 --timeout 900 -- 'bash tools/next_gpu_run_multi.sh'
# fp32 weak scaling at 1/2/4/8, strong scaling on the full 65536-row grid where it fits (fp64: 4 and 8 GPUs with K = 6;
# fp32: 2, 4, 8), the slab parity check incl. frame modes 1/2, the sharded drop-in Solver (shard.py) against one GPU, and the CPU baseline at N = 1024..8192 on the box's host.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --no-cpu --no-e2e --steps 120 --warmup 12"
port=29611
run() { name=$1; n=$2; shift 2; port=$((port + 1));
        if [ "$n" = 1 ]; then timeout 240 python $B --gpus 1 "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err
        else timeout 240 $TR --nproc-per-node $n --master-port $port $B --gpus $n "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err; fi; }
FDTD_SLAB_CHECK_FRAME=1 FDTD_SLAB_CHECK_DEEP=1 timeout 600 $TR --nproc-per-node 4 --master-port $port tools/slab_check.py > gpurun_out/multi_slab_check.log 2>&1
port=$((port + 1)); timeout 300 $TR --nproc-per-node 4 --master-port $port tools/solver_sharded_check.py > gpurun_out/multi_solver_sharded.log 2>&1
for n in 1 2 4 8; do run weak_f64_n$n $n; run weak_f32_n$n $n --dtype f32; done
for n in 4 8; do run strong_f64_n$n $n --total-rows 65536; done
for n in 2 4 8; do run deep_f64_n$n $n --halo-depth 7; run deep_window_f64_n$n $n --halo-depth 7 --frame-mode 1; done   # one exchange per K = 6 group
for n in 2 4 8; do run strong_f32_n$n $n --total-rows 65536 --dtype f32; done
timeout 300 python tools/cpu_baseline_sizes.py > gpurun_out/multi_cpu_baseline_sizes.json 2> gpurun_out/multi_cpu_baseline_sizes.err
tail -2 gpurun_out/multi_slab_check.log; tail -3 gpurun_out/multi_solver_sharded.log; for f in gpurun_out/multi_*_n*.json; do echo "$f: $(cut -c1-160 $f)"; done
