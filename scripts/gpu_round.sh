#!/bin/bash
# measurement pass: GPU tests, smoke, bench (own + reference arm), ncu launch list + full capture of the dominant kernel
set -u
mkdir -p gpurun_out
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/pytest_gpu.log
echo "== bench reference arm"; timeout 900 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.json 2>gpurun_out/bench_reference.err; echo "rc=$?"; cut -c1-200 gpurun_out/bench_reference.json
echo "== bench"; timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches.csv python scripts/prof_encode.py --warmup 1 --steps 2 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bomp_tc_kernel -s 1 -c 1 -o gpurun_out/prof_tc -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
