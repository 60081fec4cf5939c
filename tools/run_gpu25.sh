#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -8
VQB_KERNEL_TABLE=2 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32_x.json 2> gpurun_out/bench_tbl_x.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_b32_x.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])"
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/step_profile_b32.txt 2>&1
grep -E "STEP|gn_bwd|gn_apply|conv_gemm_kernel|wgrad_gemm|colsum|wgrad_reduce|GPU span" gpurun_out/step_profile_b32.txt | cut -c1-60,100-200 | head -20
