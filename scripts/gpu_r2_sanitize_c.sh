#!/bin/bash
# round 2: racecheck of the late kernels, memcheck of the earlier round-2 kernels (sweeps, exact K-SVD, omp, screen mode, ODL GEMM)
set -u
mkdir -p gpurun_out
timeout -s KILL 35 compute-sanitizer --tool racecheck --print-limit 5 python scripts/sanitize_r2b.py > gpurun_out/r2b_sanitize_racecheck.log 2>&1; echo "rc=$?"
grep -E "^ok|SUMMARY|Error|Hazard" gpurun_out/r2b_sanitize_racecheck.log | cut -c1-220 | head -12
timeout -s KILL 45 compute-sanitizer --tool memcheck --print-limit 5 python scripts/sanitize_r2.py > gpurun_out/r2_sanitize_memcheck.log 2>&1; echo "rc=$?"
grep -E "^ok|SUMMARY|Error|Invalid" gpurun_out/r2_sanitize_memcheck.log | cut -c1-220 | head -16
