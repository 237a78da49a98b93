#!/bin/bash
# First GPU call of the next round (about 4 minutes of box time; every leg has its own timeout and log under gpurun_out/):
# the kernels that were built and CPU-pinned after round 1's GPU budget had run out get their first run, and the two
# new kernels of round 1 get their ncu profiles.
#   gpurun --timeout 600 -- 'bash tools/next_gpu_run.sh'
mkdir -p gpurun_out
export FDTD_TEST_FRAME_WINDOW=1 FDTD_TEST_RESIDENT_DEPTH=1 FDTD_TEST_KSTEP_LEAN=1 FDTD_TEST_DEEP_HALO=1
# 1. parity of stepK_lean, of frame_window (frame modes 1/2), of resident_window (2..8 calls per barrier), of deep-halo slabs
#    (several slab contexts on this one GPU) and of Solver.record_field
timeout 260 python -m pytest tests/test_gpu_kstep_lean.py tests/test_gpu_frame_window.py tests/test_gpu_resident.py tests/test_gpu_deep_halo.py tests/test_zz_gpu_record_field.py -q -p no:cacheprovider --durations=8 \
    > gpurun_out/next_new_kernels_tests.log 2>&1; echo "rc=$?" >> gpurun_out/next_new_kernels_tests.log
# 2. what they buy: the C5 slab with the frame as masked passes / next to / behind the K-step kernel
for fm in 0 1 2; do
    timeout 60 python bench.py --no-cpu --no-e2e --steps 120 --warmup 12 --frame-mode $fm > gpurun_out/next_bench_frame_mode$fm.json \
        2> gpurun_out/next_bench_frame_mode$fm.err
done
# 2b. the lean K-step kernel (about 30 % fewer issue slots per row), alone and with the frame in one launch
for v in "--kstep-variant 1" "--kstep-variant 1 --frame-mode 1" "--kstep-variant 1 --temporal 8" "--kstep-variant 1 --temporal 4"; do
    n=$(echo $v | tr -d ' -'); timeout 60 python bench.py --no-cpu --no-e2e --steps 120 --warmup 12 $v > gpurun_out/next_bench_$n.json 2> gpurun_out/next_bench_$n.err
done
# 2b'. the two build knobs of stepK_lean (side-by-side libraries, ~30 s of nvcc each unless they travelled with the snapshot):
#      two CTAs per SM at K = 6 (255 registers, no spills) and the row loop unrolled by ten (delay line renamed too)
FDTD_LIB_SUFFIX=_mb2 FDTD_NVCC_EXTRA=-DFDTD_LEAN_MIN_BLOCKS_K6=2 timeout 150 python bench.py --no-cpu --no-e2e --steps 120 --warmup 12 --kstep-variant 1 \
    > gpurun_out/next_bench_lean_minblocks2.json 2> gpurun_out/next_bench_lean_minblocks2.err
FDTD_LIB_SUFFIX=_u10 FDTD_NVCC_EXTRA=-DFDTD_LEAN_UNROLL=10 timeout 150 python bench.py --no-cpu --no-e2e --steps 120 --warmup 12 --kstep-variant 1 \
    > gpurun_out/next_bench_lean_unroll10.json 2> gpurun_out/next_bench_lean_unroll10.err
# 2c. SURVEY.md 8(d) "with and without snapshot D2H": four 4.3 GB snapshots streamed to the host during 240 steps
timeout 120 python bench.py --no-cpu --no-e2e --steps 240 --warmup 6 --snapshot-every 60 > gpurun_out/next_bench_snapshots.json 2> gpurun_out/next_bench_snapshots.err
# 3. the script configs with 1 / 2 / 4 calls per grid barrier
for d in 1 2 4; do
    FDTD_RESIDENT_DEPTH=$d timeout 40 python tools/resident_times.py waveguide lens tutorial doubleslit \
        > gpurun_out/next_resident_depth$d.jsonl 2> gpurun_out/next_resident_depth$d.err
done
# 4. ncu: launch list + full capture of step_resident and frame_window (only after the plain runs above exited)
timeout 120 ncu --set full --clock-control none --import-source on -k regex:step_resident -c 1 -o gpurun_out/next_step_resident \
    python tools/resident_times.py waveguide > gpurun_out/next_ncu_resident.log 2>&1
timeout 180 ncu --set full --clock-control none --import-source on -k regex:frame_window -c 1 -o gpurun_out/next_frame_window \
    python bench.py --no-cpu --no-e2e --steps 12 --warmup 6 --frame-mode 1 > gpurun_out/next_ncu_frame_window.log 2>&1
tail -3 gpurun_out/next_new_kernels_tests.log; cat gpurun_out/next_bench_*.json | cut -c1-300; cat gpurun_out/next_resident_depth*.jsonl | cut -c1-400
