#!/bin/bash
# round 2: full single-GPU pass: all GPU tests, smoke, bench (both arms)
set -u
mkdir -p gpurun_out
echo "== all gpu tests"; timeout 1500 python -m pytest tests -m gpu -x -q --timeout 400 > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?"; tail -6 gpurun_out/r2_pytest_all.log
echo "== smoke"; timeout 300 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== odl time"; timeout 120 python scripts/odl_time.py 2>&1 | tail -6
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])); print(json.dumps(d['extras'], indent=1)[:3500])"; tail -3 gpurun_out/r2_bench_n1.err
