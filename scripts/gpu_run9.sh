#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== learners"; timeout 900 python scripts/bench_learners.py --skip-odl > gpurun_out/learners.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/learners.log
echo "== ncu sweep"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ksvd_sweep_kernel -c 1 -o gpurun_out/prof_r9_sweep -f python scripts/prof_sweep.py --signals 500000 > gpurun_out/ncu_sweep.log 2>&1; echo "rc=$?"; tail -2 gpurun_out/ncu_sweep.log
