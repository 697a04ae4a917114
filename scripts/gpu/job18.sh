#!/bin/bash
# k_visc_quad (variant 13) against k_visc_team (variant 9): parity tests, check at nel 32, sweep at nel 73, ncu
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "visc_team_kernel or record_layout or record_order" > gpurun_out/j18_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j18_pytest.log
tail -5 gpurun_out/j18_pytest.log
timeout 600 python scripts/gpu/sweep.py --nel 32 --visc --variants=-1,9,13 --dss 0 --check > gpurun_out/j18_check_visc.log 2>&1; cat gpurun_out/j18_check_visc.log
timeout 900 python scripts/gpu/sweep.py --nel 73 --visc --variants=9,13 --dss 1 > gpurun_out/j18_sweep_visc.log 2>&1; cat gpurun_out/j18_sweep_visc.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_visc_quad -s 2 -c 1 -o gpurun_out/j18_prof_visc_quad python scripts/gpu/sweep.py --nel 32 --visc --variants=13 --steps 2 > gpurun_out/j18_ncu_visc.log 2>&1
tail -2 gpurun_out/j18_ncu_visc.log
