#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scripts/fused_check.py --set single --limit 60 > gpurun_out/fused_single.log 2>&1; echo "single rc=$?"
timeout 300 python scripts/fused_check.py --set pair --limit 60 > gpurun_out/fused_pair.log 2>&1; echo "pair rc=$?"
( timeout 120 python scripts/quick_time.py; TC_k=10 timeout 120 python scripts/quick_time.py ) > gpurun_out/variants.log 2>&1
grep -h -v "^   (" gpurun_out/fused_single.log gpurun_out/fused_pair.log gpurun_out/variants.log
