#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=|SKIP" gpurun_out/pytest_gpu.log | tail -12
VQB_PROFILE=1 timeout 300 python tools/step_bench.py 32 128 > gpurun_out/step_b32.log 2>&1; grep -E "STEP|vqb::|aten::add|aten::fill|AdamW" gpurun_out/step_b32.log | cut -c1-220 | head -24
VQB_GN_STATS_FUSION=0 timeout 300 python tools/step_bench.py 32 128 > gpurun_out/step_b32_nofuse.log 2>&1; grep -E "STEP" gpurun_out/step_b32_nofuse.log
