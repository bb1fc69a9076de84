#!/bin/bash
# round 2, N GPUs (N > 2 exercises the peer loops of the sweep mailboxes): sharded-sweep parity (dist_check), then the
# bench with the parity keys in extras.  Tight timeouts and early exit: multi-GPU box time is charged N-fold.
set -u
N=${1:-4}
mkdir -p gpurun_out
nvidia-smi -L | head -8
echo "== dist_check x$N"; timeout -s KILL 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; rc=$?; echo "rc=$rc"; tail -2 gpurun_out/r2_dist_check_n$N.log | cut -c1-1800
if [ $rc -ne 0 ]; then grep -v "^\*\*\*\|OMP_NUM" gpurun_out/r2_dist_check_n$N.log | head -30; exit 1; fi
echo "== bench x$N"; timeout -s KILL 240 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_n$N.json 2> gpurun_out/r2_bench_n$N.err; echo "rc=$?"; tail -1 gpurun_out/r2_bench_n$N.json | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('value %.4g  ms/step %.3f  e2e %.4g' % (d['value'], d['ms_per_step'], d['e2e']['value'])); print(json.dumps(d['extras'], indent=1)[:5000])"; tail -3 gpurun_out/r2_bench_n$N.err
