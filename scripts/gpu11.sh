set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pencil or reciprocal" > gpurun_out/pytest_gpu_pencil.log 2>&1; echo "pytest pencil rc=$?"; tail -3 gpurun_out/pytest_gpu_pencil.log
run() { timeout 600 python bench.py --steps 20 --dss-mode 1 --no-cpu --no-e2e "$@" > gpurun_out/bench_tmp.log 2>&1; echo "bench $* rc=$?"; tail -1 gpurun_out/bench_tmp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'], d['config']['fused_stage_ms_per_step'])"; }
for v in $VARIANTS; do run --elem-kernel $v; cp gpurun_out/bench_tmp.log gpurun_out/bench_v$v.log; done
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed,l1tex__data_pipe_lsu_wavefronts_mem_shared.sum,l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum,sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.per_cycle_active,launch__registers_per_thread,launch__grid_size,launch__block_size,smsp__inst_executed.sum,lts__t_sector_hit_rate.pct,sm__cycles_elapsed.max
for v in $NCUV; do
timeout 600 ncu --metrics $M --clock-control none -k regex:k_elem -s 3 -c 1 --csv --log-file gpurun_out/ncu_metrics_v$v.csv python bench.py --nel 32 --steps 2 --warmup 3 --no-cpu --no-e2e --dss-mode 1 --elem-kernel $v > gpurun_out/ncu_m$v.log 2>&1; echo "ncu metrics v$v rc=$?"
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem -s 3 -c 1 -o gpurun_out/prof_elem_v$v python bench.py --nel 32 --steps 2 --warmup 3 --no-cpu --no-e2e --dss-mode 1 --elem-kernel $v > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
done
