#!/bin/bash
# quick single-GPU bench (headline only)
set -u
mkdir -p gpurun_out
timeout -s KILL 300 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline > gpurun_out/r2_bench_quick.json 2> gpurun_out/r2_bench_quick.err; echo "rc=$?"
tail -1 gpurun_out/r2_bench_quick.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  kernel %.3f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['roofline']['frac'])); print(d['clocks']); print(d.get('parity'))"; tail -3 gpurun_out/r2_bench_quick.err
