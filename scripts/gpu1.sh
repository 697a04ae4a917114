set -x
mkdir -p gpurun_out
nvidia-smi > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?"
tail -15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?"; tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --nel 32 --steps 20 > gpurun_out/bench_nel32.log 2>&1; echo "bench32 rc=$?"; tail -3 gpurun_out/bench_nel32.log
timeout 600 python bench.py --nel 32 --steps 20 --dss-mode 1 --no-cpu > gpurun_out/bench_nel32_dss1.log 2>&1; echo "bench32 dss1 rc=$?"; tail -3 gpurun_out/bench_nel32_dss1.log
timeout 1200 python bench.py --steps 20 > gpurun_out/bench_default.log 2>&1; echo "bench rc=$?"; tail -3 gpurun_out/bench_default.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_nel32.csv python bench.py --nel 32 --steps 2 --warmup 3 --no-cpu --no-e2e > gpurun_out/ncu_b.log 2>&1; echo "ncu rc=$?"
