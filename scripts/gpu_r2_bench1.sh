#!/bin/bash
# full single-GPU bench with phase marks
set -u
mkdir -p gpurun_out
timeout -s KILL 420 python bench.py --steps 10 --warmup 3 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err; echo "rc=$?"
grep "bench rank" gpurun_out/r2_bench_n1.err
tail -1 gpurun_out/r2_bench_n1.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  kernel %.3f e2e %.4g frac %.3f' % (d['value'], d['ms_per_step'], d['roofline']['kernel_ms_per_step'], d['e2e']['value'], d['roofline']['frac'])); print(json.dumps(d['extras'])[:3000])"
