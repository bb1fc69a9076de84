#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== gemm check"; timeout 300 python scripts/check_gemm.py > gpurun_out/check_gemm.log 2>&1; echo "rc=$?"; cat gpurun_out/check_gemm.log | tail -20
echo "== pytest gpu"; timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; cat gpurun_out/bench.json; tail -5 gpurun_out/bench.err
for mb in 16 32 128; do echo "== bench chunk ${mb}MB"; LYS_CHUNK_MB=$mb timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline 2>&1 | python -c "import sys,json; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'])"; done
echo "== ncu launches"; timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/launches_r3.csv python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_launches.log 2>&1; echo "rc=$?"
echo "== ncu full gemm"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:corr_gemm_tc -s 33 -c 1 -o gpurun_out/prof_r3_gemm -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_full.log 2>&1; echo "rc=$?"
ls -la gpurun_out | head -20
