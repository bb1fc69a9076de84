#!/bin/bash
mkdir -p gpurun_out
export LYS_TC_VER=3
timeout 300 python scripts/fused_check.py --set single --limit 60 > gpurun_out/fused_single.log 2>&1; echo "single rc=$?"
timeout 300 python scripts/fused_check.py --set pair --limit 60 > gpurun_out/fused_pair.log 2>&1; echo "pair rc=$?"
grep -h -v "^   (" gpurun_out/fused_single.log gpurun_out/fused_pair.log
