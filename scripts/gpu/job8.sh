#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py -x -q -k "visc_team or one_rhs_3d or config2 or atomics_dss" > gpurun_out/j8_pytest.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j8_pytest.log
tail -6 gpurun_out/j8_pytest.log
timeout 900 python scripts/gpu/sweep.py --nel 73 --visc --variants=9,-1 --dss 1 > gpurun_out/j8_sweep_visc.log 2>&1
cat gpurun_out/j8_sweep_visc.log
timeout 600 python scripts/gpu/sweep.py --nel 32 --visc --variants=-1,9 --dss 0 --check > gpurun_out/j8_check_visc.log 2>&1
cat gpurun_out/j8_check_visc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_visc_team -s 2 -c 1 -o gpurun_out/j8_prof_visc python scripts/gpu/sweep.py --nel 32 --visc --variants=9 --steps 2 > gpurun_out/j8_ncu_visc.log 2>&1
tail -2 gpurun_out/j8_ncu_visc.log
