#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
tail -n 60 gpurun_out/pytest_gpu.log
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 8 128 > gpurun_out/step_bench.log 2>&1
echo "step_bench exit=$?" >> gpurun_out/step_bench.log
tail -n 70 gpurun_out/step_bench.log
