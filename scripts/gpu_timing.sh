#!/bin/bash
mkdir -p gpurun_out
( LYS_TC_TIMING=1 TC_k=10 timeout 200 python scripts/tc_timing.py ) > gpurun_out/tc_timing_k10.log 2>&1
cat gpurun_out/tc_timing_k10.log
TC_k=10 timeout 100 python scripts/quick_time.py
