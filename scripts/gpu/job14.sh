#!/bin/bash
# fused stage update as one-node-per-thread launch, exit budget of the interior launch, C3 / C4 at their stated sizes
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/j14_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j14_pytest.log
tail -5 gpurun_out/j14_pytest.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 100 > gpurun_out/j14_bench_default.json 2> gpurun_out/j14_bench_default.err; cat gpurun_out/j14_bench_default.json
timeout 600 python bench.py --nop 7 --nel 41 --no-cpu --no-e2e --steps 30 > gpurun_out/j14_bench_nop7.json 2> gpurun_out/j14_bench_nop7.err; cat gpurun_out/j14_bench_nop7.json
