#!/bin/bash
mkdir -p gpurun_out
( timeout 120 python scripts/quick_time.py; for v in 400 900 1500; do LYSSA_B200_LIB=$PWD/lyssandra_b200/liblyssa_var_$v.so timeout 120 python scripts/quick_time.py; done ) > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
echo "== sanitizer (memcheck + racecheck on small fused shapes)"
timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_small.py > gpurun_out/sanitize_memcheck.log 2>&1; tail -3 gpurun_out/sanitize_memcheck.log
timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_small.py > gpurun_out/sanitize_racecheck.log 2>&1; tail -3 gpurun_out/sanitize_racecheck.log
