#!/bin/bash
# round 2: ncu evidence — launch lists + full captures of the dominant kernels
set -u
mkdir -p gpurun_out
echo "== launch list: encode (bench step)"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2_encode_launches.csv python scripts/prof_encode.py --warmup 1 --steps 2 > gpurun_out/ncu_l1.log 2>&1; echo "rc=$?"
echo "== launch list: K-SVD sweep"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/r2_sweep_launches.csv python scripts/prof_sweep.py > gpurun_out/ncu_l2.log 2>&1; echo "rc=$?"
echo "== launch list: ODL"; timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 120 --csv --log-file gpurun_out/r2_odl_launches.csv python scripts/prof_odl.py > gpurun_out/ncu_l3.log 2>&1; echo "rc=$?"
echo "== full: bomp_tc_kernel"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:bomp_tc_kernel -s 1 -c 1 -o gpurun_out/r2_prof_tc -f python scripts/prof_encode.py --warmup 1 --steps 1 > gpurun_out/ncu_f1.log 2>&1; echo "rc=$?"
echo "== full: ksvd_sweep_kernel"; timeout 900 ncu --set full --clock-control none --import-source on -k regex:ksvd_sweep_kernel -c 1 -o gpurun_out/r2_prof_sweep -f python scripts/prof_sweep.py > gpurun_out/ncu_f2.log 2>&1; echo "rc=$?"; tail -3 gpurun_out/ncu_f2.log
echo "== full: ODL kernels"; timeout 600 ncu --set full --clock-control none --import-source on -k regex:"da_gemm_tc_kernel|odl_accumulate_kernel|odl_update_kernel" -s 3 -c 3 -o gpurun_out/r2_prof_odl -f python scripts/prof_odl.py > gpurun_out/ncu_f3.log 2>&1; echo "rc=$?"
ls -la gpurun_out/*.ncu-rep
