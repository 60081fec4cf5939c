#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
VQB_KERNEL_TABLE=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32_halo.json 2> gpurun_out/bench_tbl_halo.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_b32_halo.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline'], d['roofline_wgrad'])"
