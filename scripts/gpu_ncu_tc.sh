#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:bomp_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_tc.log 2>&1; echo "ncu rc=$?"
