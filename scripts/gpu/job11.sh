#!/bin/bash
# round-2 validation of the whole tree on one B200: GPU test suite, smoke, bench lines of every BASELINE configuration at N = 1
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/j11_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j11_pytest.log
tail -5 gpurun_out/j11_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j11_smoke.log 2>&1; tail -2 gpurun_out/j11_smoke.log
timeout 600 python bench.py > gpurun_out/j11_bench_default.json 2> gpurun_out/j11_bench_default.err; cat gpurun_out/j11_bench_default.json
timeout 600 python bench.py --visc --no-cpu --no-e2e --steps 30 > gpurun_out/j11_bench_visc.json 2> gpurun_out/j11_bench_visc.err; cat gpurun_out/j11_bench_visc.json
timeout 600 python bench.py --nop 7 --nel 41 --no-cpu --no-e2e --steps 30 > gpurun_out/j11_bench_nop7.json 2> gpurun_out/j11_bench_nop7.err; cat gpurun_out/j11_bench_nop7.json
for c in c2 c3 c4; do
timeout 600 python bench.py --config $c --no-cpu --steps 100 > gpurun_out/j11_bench_$c.json 2> gpurun_out/j11_bench_$c.err; cat gpurun_out/j11_bench_$c.json
done
