#!/bin/bash
# source-level captures of the viscous pass and of the team kernel after the prefetch fix; counters of the SGS closure kernel
mkdir -p gpurun_out
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_visc_quad -s 2 -c 1 -f -o gpurun_out/j35_visc_quad_nel64 python scripts/gpu/sweep.py --nel 64 --visc --variants=13 --steps 2 > gpurun_out/j35_a.log 2>&1
timeout 200 ncu --set full --import-source on --clock-control none -k regex:k_elem_team -s 2 -c 1 -f -o gpurun_out/j35_team_nel48 python scripts/gpu/sweep.py --nel 48 --variants=9 --steps 2 > gpurun_out/j35_b.log 2>&1
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,lts__t_sector_hit_rate.pct,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_wait_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,launch__registers_per_thread,launch__occupancy_limit_registers,launch__occupancy_limit_shared_mem,sm__warps_active.avg.pct_of_peak_sustained_active
cat > gpurun_out/j35_sgs.py <<'PY'
import sys
sys.argv = ["bench.py", "--visc", "--visc-model", "VREM", "--nel", "32", "--steps", "3", "--no-cpu", "--no-e2e", "--graph", "0"]
import runpy
runpy.run_path("bench.py", run_name="__main__")
PY
timeout 200 ncu --metrics $M --clock-control none -k regex:k_elem_node -s 3 -c 1 --csv --log-file gpurun_out/j35_ncu_sgs_vrem_nel32.csv python gpurun_out/j35_sgs.py > gpurun_out/j35_c.log 2>&1
ls -la gpurun_out/j35*; tail -2 gpurun_out/j35_a.log gpurun_out/j35_b.log; grep -v "^==" gpurun_out/j35_ncu_sgs_vrem_nel32.csv | cut -d, -f13- | head -30
