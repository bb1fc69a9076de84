#!/bin/bash
# round 2 (late kernels): compute-sanitizer memcheck over the fused 'thresh' scan, the dense-row writer, the ODL bitmap kernel
set -u
mkdir -p gpurun_out
timeout -s KILL 110 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_r2b.py > gpurun_out/r2b_sanitize_memcheck.log 2>&1; echo "rc=$?"
grep -E "^ok|SUMMARY|Error|Invalid" gpurun_out/r2b_sanitize_memcheck.log | cut -c1-220 | head -16
