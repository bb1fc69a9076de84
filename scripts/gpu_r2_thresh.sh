#!/bin/bash
# round 2: two-pass 'thresh' scan in the fused kernel + double-buffered dense-row writer: timing first (a hang shows
# within a minute), then the parity tests.  Tight timeouts: a dead-locked kernel must not burn the GPU budget.
set -u
mkdir -p gpurun_out
echo "== thresh time"; timeout -s KILL 90 python scripts/thresh_time.py 2>&1 | tail -8 || { echo "TIMING HUNG/FAILED"; exit 1; }
echo "== thresh + encode tests"; timeout -s KILL 400 python -m pytest tests/test_gpu_thresh.py tests/test_gpu_encode.py -x -q --timeout 120 2>&1 | tail -8
