#!/bin/bash
mkdir -p gpurun_out
JX_LIB=$PWD/jexpresso_b200/lib_min/libjexrhs.so timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 9,10 --dss 1 > gpurun_out/j3_sweep_min.log 2>&1
cat gpurun_out/j3_sweep_min.log
JX_LIB=$PWD/jexpresso_b200/lib_fake/libjexrhs.so timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 10 --dss 1 > gpurun_out/j3_sweep_fake.log 2>&1
cat gpurun_out/j3_sweep_fake.log
