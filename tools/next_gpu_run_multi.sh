#!/bin/bash
# Multi-GPU legs still missing from SURVEY.md 8(d) after round 1.  An N-GPU call is charged N x its box time, so every
# leg runs in the call that has exactly the GPUs it needs:
#   gpurun --gpus 2 --timeout 600 -- 'bash tools/next_gpu_run_multi.sh 2'    # parity checks (2 ranks) + the N = 2 bench legs
#   gpurun --gpus 4 --timeout 420 -- 'bash tools/next_gpu_run_multi.sh 4'    # the N = 4 bench legs
#   gpurun --gpus 8 --timeout 420 -- 'bash tools/next_gpu_run_multi.sh 8'    # the N = 8 bench legs
# Legs per N: fp64 and fp32 weak scaling (8192 rows per GPU), strong scaling on the full 65536-row grid where it fits
# (fp64 from 4 GPUs at K = 6, fp32 from 2), and deep halos (one exchange per K = 6 group, the frame in frame_window) with
# the window launch behind and next to the K-step kernel.  The N = 2 call also runs the slab parity check incl. frame
# modes 1/2 and deep halos, the sharded drop-in Solver against one GPU, the same-box N = 1 reference lines and the CPU
# baseline at N = 1024..8192 on the box's host.
N=${1:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
B="bench.py --no-cpu --no-e2e --steps 120 --warmup 12"
port=29611
run() { name=$1; n=$2; shift 2; port=$((port + 1));
        if [ "$n" = 1 ]; then timeout 200 python $B --gpus 1 "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err
        else timeout 200 $TR --nproc-per-node $n --master-port $port $B --gpus $n "$@" > gpurun_out/multi_$name.json 2> gpurun_out/multi_$name.err; fi; }
if [ "$N" = 2 ]; then
    FDTD_SLAB_CHECK_FRAME=1 FDTD_SLAB_CHECK_DEEP=1 timeout 420 $TR --nproc-per-node 2 --master-port $port tools/slab_check.py > gpurun_out/multi_slab_check.log 2>&1
    port=$((port + 1)); timeout 240 $TR --nproc-per-node 2 --master-port $port tests/sharded_solver_check.py > gpurun_out/multi_solver_sharded.log 2>&1
    run weak_f64_n1 1; run weak_f32_n1 1 --dtype f32                  # the same-box N = 1 lines the efficiencies refer to
    timeout 300 python tools/cpu_baseline_sizes.py > gpurun_out/multi_cpu_baseline_sizes.json 2> gpurun_out/multi_cpu_baseline_sizes.err
fi
run weak_f64_n$N $N
run weak_f32_n$N $N --dtype f32
if [ "$N" -ge 4 ]; then run strong_f64_n$N $N --total-rows 65536; fi
run strong_f32_n$N $N --total-rows 65536 --dtype f32
run deep_f64_n$N $N --halo-depth 7                                   # window launch behind the K-step kernel
run deep_window_f64_n$N $N --halo-depth 7 --frame-mode 1              # ... and on the edge stream
tail -2 gpurun_out/multi_slab_check.log 2>/dev/null; tail -3 gpurun_out/multi_solver_sharded.log 2>/dev/null
for f in gpurun_out/multi_*_n$N.json; do echo "$f: $(cut -c1-160 $f)"; done
