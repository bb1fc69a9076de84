#!/bin/bash
set -u
mkdir -p gpurun_out
for i in 1 2; do
echo "== sweep time rowmap1 (default)"; timeout 120 python scripts/sweep_time.py 2>&1 | tail -1
echo "== sweep time rowmap0"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_rowmap0.so timeout 120 python scripts/sweep_time.py 2>&1 | tail -1
done
echo "== odl time"; timeout 120 python scripts/odl_time.py 2>&1 | tail -7
echo "== odl tests"; timeout 600 python -m pytest tests/test_gpu_learners.py -x -q --timeout 300 -k "odl" 2>&1 | tail -5
