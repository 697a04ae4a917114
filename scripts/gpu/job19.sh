#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "visc_team_kernel" > gpurun_out/j19_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j19_pytest.log
tail -3 gpurun_out/j19_pytest.log
timeout 900 python scripts/gpu/sweep.py --nel 73 --visc --variants=13 --dss 1 > gpurun_out/j19_sweep_visc.log 2>&1; cat gpurun_out/j19_sweep_visc.log
M=gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__t_requests_pipe_lsu_mem_local_op_ld.sum
timeout 600 ncu --metrics $M --clock-control none -k regex:k_visc_quad -s 2 -c 1 --csv --log-file gpurun_out/j19_ncu_vq.csv python scripts/gpu/sweep.py --nel 32 --visc --variants=13 --steps 2 > gpurun_out/j19_ncu_vq.log 2>&1
grep -v "^==" gpurun_out/j19_ncu_vq.csv | cut -d, -f13-
