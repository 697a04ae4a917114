#!/bin/bash
mkdir -p gpurun_out
export JX_LIB=$PWD/jexpresso_b200/lib_min/libjexrhs.so
timeout 600 python scripts/gpu/sweep.py --nel 40 --variants 9,10,11 --dss 0 --check > gpurun_out/j6_check.log 2>&1
cat gpurun_out/j6_check.log
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 9,11 --dss 1 > gpurun_out/j6_sweep.log 2>&1
cat gpurun_out/j6_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_team2 -s 3 -c 1 -o gpurun_out/j6_prof_t3 python scripts/gpu/sweep.py --nel 32 --variants 11 --steps 2 > gpurun_out/j6_ncu_t3.log 2>&1
tail -2 gpurun_out/j6_ncu_t3.log
