set -x
mkdir -p gpurun_out
timeout -k 10 600 python -m pytest tests -m gpu -x -q -k "pencil or atomics or periodic or config2" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sel.log
timeout -k 10 400 python bench.py --no-cpu --no-e2e > gpurun_out/bench_tmp.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_tmp.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'], d['config']['eager_ms_per_step'], d['config']['fused_stage_ms_per_step'])"
