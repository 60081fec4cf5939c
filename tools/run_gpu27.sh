#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py fat 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback|fat_conv" | head
timeout 300 python tools/gpu_probe.py tinybench 2>&1 | grep -E "BENCH|Error|error|Traceback"
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
