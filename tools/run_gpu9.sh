#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
for B in 32 48 64; do timeout 300 python tools/step_bench.py $B 128 > gpurun_out/step_b$B.log 2>&1; grep STEP gpurun_out/step_b$B.log; done
timeout 300 python tools/step_bench.py 16 128 gan > gpurun_out/step_b16_gan.log 2>&1; grep STEP gpurun_out/step_b16_gan.log
