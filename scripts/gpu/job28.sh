#!/bin/bash
# final tree of round 2: GPU suite, smoke, the default bench line exactly as the driver runs it
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/j28_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j28_pytest.log
tail -4 gpurun_out/j28_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j28_smoke.log 2>&1; tail -1 gpurun_out/j28_smoke.log
timeout 900 python bench.py > gpurun_out/j28_bench_default.json 2> gpurun_out/j28_bench_default.err; cut -c1-300 gpurun_out/j28_bench_default.json
