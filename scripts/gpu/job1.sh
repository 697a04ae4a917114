#!/bin/bash
# round-2 job 1: functor parity after the total-energy fix + gather micro-benchmark
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/j1_smi.txt 2>&1
timeout 900 python -m pytest tests/test_zzz_functors_gpu.py -x -q > gpurun_out/j1_pytest_functors.log 2>&1
echo "pytest rc=$?" >> gpurun_out/j1_pytest_functors.log
cd scripts/micro
( timeout -s KILL 120 ./gather 48 1 -1 ) > ../../gpurun_out/j1_gather_box1.log 2>&1
( timeout -s KILL 60 ./gather 48 4 3 ) > ../../gpurun_out/j1_gather_box4.log 2>&1
( timeout -s KILL 300 ncu --metrics gpu__time_duration.sum,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,lts__t_sectors_srcunit_tex_op_read.sum,dram__bytes_read.sum,smsp__inst_executed.sum --clock-control none --csv --log-file ../../gpurun_out/j1_gather_ncu.csv ./gather 48 1 -1 ) > ../../gpurun_out/j1_gather_ncu.log 2>&1
cd ../..
tail -3 gpurun_out/j1_pytest_functors.log; cat gpurun_out/j1_gather_box1.log gpurun_out/j1_gather_box4.log
