set -x
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -x -q -k "pencil or reciprocal" > gpurun_out/pytest_gpu_pencil.log 2>&1; echo "pytest pencil rc=$?"; tail -8 gpurun_out/pytest_gpu_pencil.log
for v in 3 4; do
timeout 600 python bench.py --steps 20 --dss-mode 1 --no-cpu --no-e2e --elem-kernel $v > gpurun_out/bench_v$v.log 2>&1; echo "bench v$v rc=$?"; tail -1 gpurun_out/bench_v$v.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'], d['config']['fused_stage_ms_per_step'], d['clocks'])"
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:k_elem -s 3 -c 1 -o gpurun_out/prof_elem_r01e python bench.py --nel 32 --steps 2 --warmup 3 --no-cpu --no-e2e --dss-mode 1 --elem-kernel 3 > gpurun_out/ncu_full.log 2>&1; echo "ncu rc=$?"
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_gpu.log
