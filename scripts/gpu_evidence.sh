set -x
mkdir -p gpurun_out
timeout -k 10 300 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/pytest_gpu.log
timeout -k 10 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/smoke.log
timeout -k 10 300 python bench.py > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -1 gpurun_out/bench_default.log | cut -c1-400
timeout -k 10 120 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_reference.log 2>&1; echo "ref rc=$?"; tail -1 gpurun_out/bench_reference.log | cut -c1-300
timeout -k 10 150 python bench.py --nop 7 --nel 41 --steps 10 --no-cpu --no-e2e > gpurun_out/bench_nop7.log 2>&1; echo "nop7 rc=$?"; tail -1 gpurun_out/bench_nop7.log | cut -c1-600
