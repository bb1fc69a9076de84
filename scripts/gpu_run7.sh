#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== learners"; timeout 900 python scripts/bench_learners.py > gpurun_out/learners.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/learners.log
echo "== ncu k10 greedy"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bomp_fast_kernel -s 1 -c 1 -o gpurun_out/prof_r7_fast_k10 -f python scripts/prof_encode.py --signals 262144 --k 10 --warmup 1 --steps 1 --sparse-only > gpurun_out/ncu_k10.log 2>&1; echo "rc=$?"
echo "== ncu sweep"; timeout 1200 ncu --set full --clock-control none --import-source on -k regex:ksvd_sweep_kernel -c 1 -o gpurun_out/prof_r7_sweep -f python scripts/prof_sweep.py --signals 500000 > gpurun_out/ncu_sweep.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_sweep.log
