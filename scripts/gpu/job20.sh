#!/bin/bash
# full validation of the tree on one B200: GPU suite, smoke, every BASELINE configuration at N = 1
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/j20_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j20_pytest.log
tail -5 gpurun_out/j20_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j20_smoke.log 2>&1; tail -2 gpurun_out/j20_smoke.log
timeout 600 python bench.py --visc --no-cpu --no-e2e --steps 50 > gpurun_out/j20_bench_visc.json 2> gpurun_out/j20_bench_visc.err; cut -c1-700 gpurun_out/j20_bench_visc.json
for c in c2 c3 c4; do
timeout 600 python bench.py --config $c --no-cpu --steps 100 > gpurun_out/j20_bench_$c.json 2> gpurun_out/j20_bench_$c.err; cut -c1-700 gpurun_out/j20_bench_$c.json
done
timeout 600 python bench.py --config c4 --overlap 4 --periodic --no-cpu --no-e2e --steps 100 > gpurun_out/j20_bench_c4_ov4.json 2> gpurun_out/j20_bench_c4_ov4.err; cut -c1-1400 gpurun_out/j20_bench_c4_ov4.json
