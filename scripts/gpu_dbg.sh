#!/bin/bash
mkdir -p gpurun_out
for d in 0 1 2 3 8 12; do echo "dbg=$d"; LYS_TC_DBG=$d timeout 120 python scripts/quick_time.py 2>&1 | grep -E "default|Error"; done > gpurun_out/dbg.log 2>&1
cat gpurun_out/dbg.log
