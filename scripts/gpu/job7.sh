#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q > gpurun_out/j7_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j7_pytest.log
tail -6 gpurun_out/j7_pytest.log
timeout 600 python scripts/gpu/sweep.py --nel 41 --nop 7 --variants 12,-1 --dss 1 > gpurun_out/j7_sweep_nop7.log 2>&1
cat gpurun_out/j7_sweep_nop7.log
timeout 600 python scripts/gpu/sweep.py --nel 20 --nop 7 --variants -1,12 --dss 0 --check > gpurun_out/j7_check_nop7.log 2>&1
cat gpurun_out/j7_check_nop7.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem_tri -s 3 -c 1 -o gpurun_out/j7_prof_tri python scripts/gpu/sweep.py --nel 20 --nop 7 --variants 12 --steps 2 > gpurun_out/j7_ncu_tri.log 2>&1
tail -2 gpurun_out/j7_ncu_tri.log
