#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -12
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/step_b32.log 2>&1; grep -E "STEP|vqb::|aten::add|AdamW|Memcpy|elementwise" gpurun_out/step_b32.log | cut -c1-230 | head -40
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench exit=$?"; cat gpurun_out/bench_b32.json; tail -5 gpurun_out/bench_b32.err
