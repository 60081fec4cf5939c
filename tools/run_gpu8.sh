#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
VQB_KERNEL_TABLE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_tbl.err; echo "bench exit=$?"; head -30 gpurun_out/bench_tbl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])"
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/step_b32.log 2>&1; grep -E "STEP|vqb::|aten::add|AdamW|elementwise|Memcpy" gpurun_out/step_b32.log | cut -c1-220 | head -34
