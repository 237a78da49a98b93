#!/bin/bash
# One bounded GPU call: first run of resident mode and of the last-pass overlap on a B200.  Every leg has its own
# timeout and log under gpurun_out/ so that a hang in one leg costs only that leg.
mkdir -p gpurun_out
export FDTD_TEST_RESIDENT=1
date +%s.%N > gpurun_out/shot_t0
timeout 45 python -m pytest tests/test_gpu_resident.py -q -p no:cacheprovider --durations=5 > gpurun_out/shot_resident_tests.log 2>&1
echo "rc=$?" >> gpurun_out/shot_resident_tests.log
timeout 25 python tools/resident_times.py > gpurun_out/shot_resident_times.jsonl 2> gpurun_out/shot_resident_times.err
echo "rc=$?" >> gpurun_out/shot_resident_times.err
timeout 25 python tools/overlap_check.py > gpurun_out/shot_overlap.json 2> gpurun_out/shot_overlap.err
echo "rc=$?" >> gpurun_out/shot_overlap.err
FDTD_RESIDENT=1 timeout 35 python -m pytest tests/test_gpu_parity.py -q -p no:cacheprovider \
    -k "not full_size and not 4096 and not mid_size and not temporal and not large_hetero" > gpurun_out/shot_parity_resident_default.log 2>&1
echo "rc=$?" >> gpurun_out/shot_parity_resident_default.log
date +%s.%N > gpurun_out/shot_t1
tail -3 gpurun_out/shot_resident_tests.log; cat gpurun_out/shot_resident_times.jsonl; cat gpurun_out/shot_overlap.json; tail -3 gpurun_out/shot_parity_resident_default.log
