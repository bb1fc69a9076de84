#!/bin/bash
# round 2: first run of the look-ahead / integer-reduction sweep kernel: parity tests, timing at cfg3
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== pytest learners"; timeout 900 python -m pytest tests/test_gpu_learners.py -x -q > gpurun_out/r2_pytest_learners.log 2>&1; echo "rc=$?"; tail -15 gpurun_out/r2_pytest_learners.log
echo "== sweep time cfg3"; timeout 300 python scripts/sweep_time.py 2>&1 | tail -3
