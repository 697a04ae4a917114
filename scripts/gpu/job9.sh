#!/bin/bash
mkdir -p gpurun_out
export JX_LIB=$PWD/jexpresso_b200/lib_min/libjexrhs.so
timeout 600 python scripts/gpu/sweep.py --nel 32 --visc --variants=-1,9 --dss 0 --check > gpurun_out/j9_check_visc.log 2>&1
cat gpurun_out/j9_check_visc.log
timeout 900 python scripts/gpu/sweep.py --nel 73 --visc --variants=9 --dss 1 > gpurun_out/j9_sweep_visc.log 2>&1
cat gpurun_out/j9_sweep_visc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_visc_team -s 2 -c 1 -o gpurun_out/j9_prof_visc python scripts/gpu/sweep.py --nel 32 --visc --variants=9 --steps 2 > gpurun_out/j9_ncu_visc.log 2>&1
tail -2 gpurun_out/j9_ncu_visc.log
