#!/bin/bash
# packed pair records + Morton element order: GPU suite, A/B of the element order, viscous bench
mkdir -p gpurun_out
timeout 2000 python -m pytest tests -m gpu -x -q > gpurun_out/j15_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j15_pytest.log
tail -5 gpurun_out/j15_pytest.log
JX_ELEM_ORDER=0 timeout 600 python scripts/gpu/sweep.py --nel 73 --variants=9 --dss 1 --steps 20 > gpurun_out/j15_sweep_order0.log 2>&1; cat gpurun_out/j15_sweep_order0.log
timeout 600 python scripts/gpu/sweep.py --nel 73 --variants=9 --dss 1 --steps 20 > gpurun_out/j15_sweep_order1.log 2>&1; cat gpurun_out/j15_sweep_order1.log
timeout 600 python bench.py --no-cpu --no-e2e --steps 100 > gpurun_out/j15_bench_default.json 2> gpurun_out/j15_bench_default.err; cat gpurun_out/j15_bench_default.json
timeout 600 python bench.py --visc --no-cpu --no-e2e --steps 30 > gpurun_out/j15_bench_visc.json 2> gpurun_out/j15_bench_visc.err; cat gpurun_out/j15_bench_visc.json
