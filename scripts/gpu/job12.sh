#!/bin/bash
# fused stage update (k_stage_fused): GPU suite + default bench line
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j12_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j12_pytest.log
tail -5 gpurun_out/j12_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 50 > gpurun_out/j12_bench_default.json 2> gpurun_out/j12_bench_default.err; cat gpurun_out/j12_bench_default.json
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants=9 --dss 1 --fused > gpurun_out/j12_sweep_fused.log 2>&1; cat gpurun_out/j12_sweep_fused.log
