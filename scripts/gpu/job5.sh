#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "team2 or interface_first" > gpurun_out/j5_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j5_pytest.log
tail -4 gpurun_out/j5_pytest.log
export JX_LIB=$PWD/jexpresso_b200/lib_min/libjexrhs.so
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 9,10 --dss 1 > gpurun_out/j5_sweep.log 2>&1
cat gpurun_out/j5_sweep.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_team2 -s 3 -c 1 -o gpurun_out/j5_prof_t2 python scripts/gpu/sweep.py --nel 32 --variants 10 --steps 2 > gpurun_out/j5_ncu_t2.log 2>&1
tail -2 gpurun_out/j5_ncu_t2.log
