#!/bin/bash
set -u
mkdir -p gpurun_out
for tool in racecheck synccheck memcheck; do
  echo "== $tool"; timeout 900 compute-sanitizer --tool $tool --print-limit 5 python scripts/sanitize_small.py > gpurun_out/sanitize_$tool.log 2>&1; echo "rc=$?"; grep -E "^ok|SUMMARY|Error" gpurun_out/sanitize_$tool.log | cut -c1-200 | head -12
done
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== learners"; timeout 900 python scripts/bench_learners.py > gpurun_out/learners.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/learners.log
