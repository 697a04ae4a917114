#!/bin/bash
# direct accumulation into the low-storage register (k_stage_direct): GPU suite, default / viscous / c4 bench lines
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/j21_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j21_pytest.log
tail -5 gpurun_out/j21_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j21_smoke.log 2>&1; tail -2 gpurun_out/j21_smoke.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 100 > gpurun_out/j21_bench_default.json 2> gpurun_out/j21_bench_default.err; cut -c1-1500 gpurun_out/j21_bench_default.json
timeout 600 python bench.py --visc --no-cpu --no-e2e --steps 50 > gpurun_out/j21_bench_visc.json 2> gpurun_out/j21_bench_visc.err; cut -c1-300 gpurun_out/j21_bench_visc.json
timeout 600 python bench.py --config c4 --no-cpu --no-e2e --steps 100 > gpurun_out/j21_bench_c4.json 2> gpurun_out/j21_bench_c4.err; cut -c1-300 gpurun_out/j21_bench_c4.json
