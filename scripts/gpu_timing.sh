#!/bin/bash
mkdir -p gpurun_out
( LYS_TC_TIMING=1 LYS_TC_SLOTS=2 timeout 200 python scripts/tc_timing.py ) > gpurun_out/tc_timing.log 2>&1
cat gpurun_out/tc_timing.log
