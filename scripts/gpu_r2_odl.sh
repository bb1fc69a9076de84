#!/bin/bash
# round 2: ODL statistics kernel with the user bitmap (no per-atom sort) + class_dict_learn: learner tests, then timing
set -u
mkdir -p gpurun_out
echo "== learner tests"; timeout -s KILL 240 python -m pytest tests/test_gpu_learners.py -x -q --timeout 100 2>&1 | tail -8
echo "== odl time"; timeout -s KILL 60 python scripts/odl_time.py 2>&1 | tail -8
