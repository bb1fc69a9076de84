#!/bin/bash
# bring-up of the fused kernel: single-CTA shapes, then CTA-pair shapes, each in its own watchdogged process
mkdir -p gpurun_out
timeout 400 python scripts/fused_check.py --set single > gpurun_out/fused_single.log 2>&1; echo "single rc=$?"
timeout 400 python scripts/fused_check.py --set pair > gpurun_out/fused_pair.log 2>&1; echo "pair rc=$?"
tail -30 gpurun_out/fused_single.log gpurun_out/fused_pair.log
