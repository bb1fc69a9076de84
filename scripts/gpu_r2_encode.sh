#!/bin/bash
# round 2: screen-mode fused encode: parity tests + timing
set -u
mkdir -p gpurun_out
echo "== quick time k=5"; timeout 300 python scripts/quick_time.py 2>&1 | tail -3
echo "== quick time k=10"; TC_k=10 timeout 300 python scripts/quick_time.py 2>&1 | tail -3
echo "== pytest encode"; timeout 1500 python -m pytest tests/test_gpu_encode.py tests/test_gpu_thresh.py -x -q > gpurun_out/r2_pytest_encode.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_pytest_encode.log
