#!/bin/bash
set -u
mkdir -p gpurun_out
echo "== bench x2"; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/bench_n2.json 2> gpurun_out/bench_n2.err; echo "rc=$?"; tail -1 gpurun_out/bench_n2.json | cut -c1-600; grep -E "Error|assert" gpurun_out/bench_n2.err | head -3
