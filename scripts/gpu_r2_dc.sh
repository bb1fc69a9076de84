#!/bin/bash
# N-GPU dist_check only (tight timeout)
set -u
N=${1:-2}
mkdir -p gpurun_out
timeout -s KILL 120 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/r2_dist_check_n$N.log 2>&1; echo "rc=$?"
grep -v "^\*\*\*\|OMP_NUM\|^$" gpurun_out/r2_dist_check_n$N.log | tail -12 | cut -c1-2500
