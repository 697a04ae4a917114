#!/bin/bash
mkdir -p gpurun_out
export JX_LIB=$PWD/jexpresso_b200/lib_min2/libjexrhs.so
timeout 600 python scripts/gpu/sweep.py --nel 32 --visc --variants=-1,9 --dss 0 --check > gpurun_out/j10_check_visc.log 2>&1
cat gpurun_out/j10_check_visc.log
timeout 900 python scripts/gpu/sweep.py --nel 73 --visc --variants=9 --dss 1 > gpurun_out/j10_sweep_visc.log 2>&1
cat gpurun_out/j10_sweep_visc.log
timeout 600 python scripts/gpu/sweep.py --nel 20 --nop 7 --variants=-1,12 --dss 0 --check > gpurun_out/j10_check_tri.log 2>&1
cat gpurun_out/j10_check_tri.log
timeout 600 python scripts/gpu/sweep.py --nel 41 --nop 7 --variants=12 --dss 1 > gpurun_out/j10_sweep_tri.log 2>&1
cat gpurun_out/j10_sweep_tri.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_tri -s 3 -c 1 -o gpurun_out/j10_prof_tri python scripts/gpu/sweep.py --nel 20 --nop 7 --variants=12 --steps 2 > gpurun_out/j10_ncu_tri.log 2>&1
tail -2 gpurun_out/j10_ncu_tri.log
