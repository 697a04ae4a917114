#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_zzz_functors_gpu.py -m gpu -x -q > gpurun_out/j29_pytest_functors.log 2>&1; echo "pytest rc=$?" >> gpurun_out/j29_pytest_functors.log
tail -15 gpurun_out/j29_pytest_functors.log
