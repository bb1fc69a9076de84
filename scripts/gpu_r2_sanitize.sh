#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in memcheck synccheck racecheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_r2.py > gpurun_out/r2_sanitize_$tool.log 2>&1; echo "rc=$?"; grep -E "^ok|SUMMARY|Error|Hazard" gpurun_out/r2_sanitize_$tool.log | cut -c1-200 | head -16
done
