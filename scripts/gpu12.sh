set -x
mkdir -p gpurun_out
timeout -k 10 400 python bench.py --no-cpu > gpurun_out/bench_graph.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_graph.log
timeout -k 10 300 python bench.py --no-cpu --no-e2e --graph 0 > gpurun_out/bench_nograph.log 2>&1; echo "bench nograph rc=$?"; tail -1 gpurun_out/bench_nograph.log | python -c "import sys,json; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], d['config']['phase_ms_per_step'])"
timeout -k 10 300 python -m pytest tests -m gpu -x -q -k "atomics or periodic or config2" > gpurun_out/pytest_sel.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_sel.log
