#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== sweep phases 2M"; LYSSA_B200_LIB=lyssandra_b200/liblyssa_b200_bringup.so timeout 120 python scripts/sweep_phases.py 2>&1 | tail -9
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?"; tail -25 gpurun_out/r2_pytest_all.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])); print(json.dumps(d['extras'], indent=1)[:3500])"; tail -3 gpurun_out/r2_bench_n1.err
