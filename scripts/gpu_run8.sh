#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/pytest_gpu.log
echo "== learners"; timeout 900 python scripts/bench_learners.py > gpurun_out/learners.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/learners.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; python -c "import sys,json; d=json.loads(open('gpurun_out/bench.json').read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; tail -5 gpurun_out/bench.err
