#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r2.csv python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bomp_fast_kernel -s 66 -c 1 -o gpurun_out/prof_r2_fast -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -20
