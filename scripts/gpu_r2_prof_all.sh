#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== encode host tests + quick bench (host densify)"; timeout 600 python -m pytest tests/test_gpu_encode.py -x -q --timeout 300 2>&1 | tail -3
timeout 600 python bench.py --steps 5 --warmup 3 --no-extras > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo "rc=$?"; tail -1 gpurun_out/r2_bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g (host dense %.1f GB/s)  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['e2e'].get('host_dense_gb_per_s_per_rank', 0), d['roofline']['frac']))"; tail -3 gpurun_out/r2_bench_quick.err
bash scripts/gpu_r2_prof.sh
