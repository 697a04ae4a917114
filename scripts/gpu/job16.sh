#!/bin/bash
# packed records + element order: GPU suite, ncu of the element kernel at full size (both element orders), launch list
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/j16_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j16_pytest.log
tail -5 gpurun_out/j16_pytest.log
M=dram__bytes_read.sum,dram__bytes_write.sum,lts__t_sector_hit_rate.pct,gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed,dram__throughput.avg.pct_of_peak_sustained_elapsed
JX_ELEM_ORDER=0 timeout 600 ncu --metrics $M --clock-control none -k regex:k_elem_team -s 3 -c 1 --csv --log-file gpurun_out/j16_ncu_order0.csv python scripts/gpu/sweep.py --nel 73 --variants=9 --steps 2 > gpurun_out/j16_ncu_order0.log 2>&1
timeout 600 ncu --metrics $M --clock-control none -k regex:k_elem_team -s 3 -c 1 --csv --log-file gpurun_out/j16_ncu_order1.csv python scripts/gpu/sweep.py --nel 73 --variants=9 --steps 2 > gpurun_out/j16_ncu_order1.log 2>&1
cat gpurun_out/j16_ncu_order0.csv gpurun_out/j16_ncu_order1.csv | grep -v "^==" | cut -d, -f5,13- 
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem_team -s 3 -c 1 -o gpurun_out/j16_prof_team_nel73 python scripts/gpu/sweep.py --nel 73 --variants=9 --steps 2 > gpurun_out/j16_ncu_full.log 2>&1
tail -2 gpurun_out/j16_ncu_full.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/j16_launches_default.csv python bench.py --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/j16_launches.log 2>&1
tail -2 gpurun_out/j16_launches.log | cut -c1-300
