#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/fused_check.py --set single --limit 60 > gpurun_out/fused_single.log 2>&1; echo "single rc=$?"
timeout 300 python scripts/fused_check.py --set pair --limit 60 > gpurun_out/fused_pair.log 2>&1; echo "pair rc=$?"
( LYS_TC_TIMING=1 timeout 200 python scripts/tc_timing.py ) > gpurun_out/tc_timing.log 2>&1
grep -h -v "^   (" gpurun_out/fused_single.log gpurun_out/fused_pair.log gpurun_out/tc_timing.log
