#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== phase timers, screen"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/tc_timing.py 2>&1 | tail -32
echo "== phase timers, split3"; TC_SPLIT3=1 LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/tc_timing.py 2>&1 | tail -32
echo "== remaining tests"; timeout 600 python -m pytest tests/test_gpu_encode.py tests/test_gpu_thresh.py -x -q -k "nan_and_inf or dictionary_scale or thresh" > gpurun_out/r2_pytest_encode2.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2_pytest_encode2.log
