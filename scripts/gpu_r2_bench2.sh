#!/bin/bash
# N-GPU bench with phase marks (every rank), tight timeout
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --extras-budget 120 > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"
grep "bench rank" gpurun_out/r2_bench_n$N.err | tail -60
tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(json.dumps(d['extras'], indent=1)[:5000])"
