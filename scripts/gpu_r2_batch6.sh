#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sweep phases 2M"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/sweep_phases.py 2>&1 | tail -9
echo "== odl time"; timeout 120 python scripts/odl_time.py 2>&1 | tail -9
echo "== learners + omp tests"; timeout 900 python -m pytest tests/test_gpu_learners.py tests/test_gpu_omp.py -x -q --timeout 300 > gpurun_out/r2_pytest_batch6.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_pytest_batch6.log
