#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py stats 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head
timeout 300 python tools/gpu_probe.py up 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head
timeout 300 python tools/gpu_probe.py halobench 2>&1 | grep -E "BENCH|Error|error|Traceback" | grep -E "128->128|64->64|Error|error" 
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
VQB_DEBUG_MODE=8192 timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
