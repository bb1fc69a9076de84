#!/bin/bash
# first GPU round trip: smoke, GPU tests, bench, ncu launch list + full capture of the top kernel
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
echo "== smoke" ; timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/smoke.log
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r1.csv python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:bomp_warp_kernel -s 66 -c 2 -o gpurun_out/prof_r1_greedy -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:sgemm_kernel -s 66 -c 1 -o gpurun_out/prof_r1_sgemm -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_full2.log 2>&1; echo "rc=$?"
ls -la gpurun_out
