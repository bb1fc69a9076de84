#!/bin/bash
# round 2, final kernels: full capture of the fused 'thresh' kernel, launch lists of the thresholding coders and of the bench step
set -u
mkdir -p gpurun_out
timeout -s KILL 150 ncu --set full --clock-control none --import-source on -k regex:bomp_tc_kernel -c 1 -o gpurun_out/r2_prof_thresh_final -f python scripts/thresh_once.py > gpurun_out/ncu_thresh.log 2>&1; echo "ncu thresh rc=$?"
timeout -s KILL 120 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_thresh_launches.csv python scripts/thresh_once.py > gpurun_out/ncu_l4.log 2>&1; echo "launch list thresh rc=$?"
timeout -s KILL 150 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_encode_launches_final.csv python scripts/prof_encode.py --warmup 1 --steps 3 > gpurun_out/ncu_l5.log 2>&1; echo "launch list encode rc=$?"
