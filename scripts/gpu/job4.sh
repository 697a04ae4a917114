#!/bin/bash
mkdir -p gpurun_out
export JX_LIB=$PWD/jexpresso_b200/lib_min/libjexrhs.so
timeout 600 python scripts/gpu/sweep.py --nel 40 --variants 9,10 --dss 0 --check > gpurun_out/j4_check.log 2>&1
cat gpurun_out/j4_check.log
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 9,10 --dss 1 > gpurun_out/j4_sweep.log 2>&1
cat gpurun_out/j4_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_team2 -s 3 -c 1 -o gpurun_out/j4_prof_t2 python scripts/gpu/sweep.py --nel 32 --variants 10 --steps 2 > gpurun_out/j4_ncu_t2.log 2>&1
tail -3 gpurun_out/j4_ncu_t2.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_team -s 3 -c 1 -o gpurun_out/j4_prof_t1 python scripts/gpu/sweep.py --nel 32 --variants 9 --steps 2 > gpurun_out/j4_ncu_t1.log 2>&1
tail -3 gpurun_out/j4_ncu_t1.log
ls -la gpurun_out/*.ncu-rep
