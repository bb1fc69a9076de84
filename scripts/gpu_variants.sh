#!/bin/bash
mkdir -p gpurun_out
for v in 1 2 3 4; do LYSSA_B200_LIB=$PWD/lyssandra_b200/liblyssa_var_$v.so timeout 120 python scripts/quick_time.py; done > gpurun_out/variants.log 2>&1
cat gpurun_out/variants.log
