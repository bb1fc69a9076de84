#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== learners"; timeout 900 python scripts/bench_learners.py > gpurun_out/learners.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/learners.log
for mb in 128 256 512; do echo "== bench chunk ${mb}MB"; LYS_CHUNK_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; done
