#!/bin/bash
# round 2, final single-GPU pass: all GPU tests, smoke, bench (own arm), tight timeouts
set -u
mkdir -p gpurun_out
echo "== all gpu tests"; timeout -s KILL 200 python -m pytest tests -m gpu -x -q --timeout 100 > gpurun_out/r2_pytest_all.log 2>&1; echo "rc=$?"; tail -4 gpurun_out/r2_pytest_all.log
echo "== smoke"; timeout -s KILL 60 python __graft_entry__.py --smoke 2>&1 | tail -2
echo "== bench"; timeout -s KILL 120 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g  frac %.3f' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['frac'])); e=d['extras']; print(e['ksvd_iteration']['ms_per_iter'], e['scspm_pipeline']['images_per_s'], e['odl_minibatch']['ms_per_minibatch'], e['odl_minibatch']['stages_ms_rank0'], e['sibling_coders']['ms_per_1M_signals'])"; grep "bench rank" gpurun_out/r2_bench_n1.err | tail -3
