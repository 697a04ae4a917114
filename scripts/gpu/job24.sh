#!/bin/bash
# viscous split fix, device-built mass + IC conditioning, k_visc_team removed: GPU suite
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -x -q > gpurun_out/j24_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j24_pytest.log
tail -5 gpurun_out/j24_pytest.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/j24_smoke.log 2>&1; tail -2 gpurun_out/j24_smoke.log
