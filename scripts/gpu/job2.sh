#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -x -q -k "team2 or one_rhs_3d_bit_exact or pencil_kernel_atomics" > gpurun_out/j2_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j2_pytest.log
tail -5 gpurun_out/j2_pytest.log
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants 9,10 --dss 1 > gpurun_out/j2_sweep_dss1.log 2>&1
cat gpurun_out/j2_sweep_dss1.log
timeout 600 python scripts/gpu/sweep.py --nel 40 --variants 0,9,10 --dss 0 --check > gpurun_out/j2_sweep_dss0.log 2>&1
cat gpurun_out/j2_sweep_dss0.log
