#!/bin/bash
set -u
mkdir -p gpurun_out
nvidia-smi -L
echo "== dist_check x2"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 scripts/dist_check.py > gpurun_out/dist_check2.log 2>&1; echo "rc=$?"; tail -5 gpurun_out/dist_check2.log
echo "== bench x2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-400; tail -3 gpurun_out/bench_n2.err
echo "== bench x2 reference arm"; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | tail -1 | cut -c1-300
