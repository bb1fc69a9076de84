#!/bin/bash
# round 2: ncu --set full of the fused 'thresh' kernel (MODE 1, two-pass scan), sparse output, 1M signals
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bomp_tc_kernel -c 1 -o gpurun_out/r2_prof_thresh -f python scripts/thresh_once.py > gpurun_out/ncu_thresh.log 2>&1; echo "ncu rc=$?"
tail -3 gpurun_out/ncu_thresh.log
