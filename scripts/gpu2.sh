set -x
mkdir -p gpurun_out
./scripts/micro/micro > gpurun_out/micro.log 2>&1; cat gpurun_out/micro.log
timeout 900 python bench.py --steps 20 --dss-mode 1 --no-cpu > gpurun_out/bench_default_dss1.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default_dss1.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'], d['config']['fused_stage_ms_per_step'])"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:k_elem -s 3 -c 1 -o gpurun_out/prof_elem_r01a python bench.py --nel 32 --steps 2 --warmup 3 --no-cpu --no-e2e --dss-mode 1 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
