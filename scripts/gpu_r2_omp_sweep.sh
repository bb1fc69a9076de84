#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sweep phases 2M"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/sweep_phases.py 2>&1 | tail -9
echo "== sweep phases 250k"; SW_N=250000 LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/sweep_phases.py 2>&1 | tail -9
echo "== omp tests"; timeout 600 python -m pytest tests/test_gpu_omp.py tests/test_gpu_encode.py -x -q --timeout 200 > gpurun_out/r2_pytest_omp.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_pytest_omp.log
