#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== quick time k=5"; timeout 120 python scripts/quick_time.py 2>&1 | tail -3
echo "== quick time k=10"; TC_k=10 timeout 120 python scripts/quick_time.py 2>&1 | tail -3
echo "== quick time K=512"; TC_K=512 timeout 120 python scripts/quick_time.py 2>&1 | tail -3
echo "== phase timers, split3"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/tc_timing.py 2>&1 | tail -14
echo "== phase timers, screen"; TC_SCREEN=1 LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/tc_timing.py 2>&1 | tail -15
echo "== tests"; timeout 900 python -m pytest tests/test_gpu_encode.py tests/test_gpu_thresh.py -x -q --timeout 300 > gpurun_out/r2_pytest_encode3.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2_pytest_encode3.log
