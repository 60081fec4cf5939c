#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -q -m gpu -s -k "attention or step" > gpurun_out/pytest_gpu2.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu2.log
grep -E "passed|failed|FAIL|Error|exit=|attn|vae_attn|cos " gpurun_out/pytest_gpu2.log | tail -30
VQB_KERNEL_TABLE=1 timeout 900 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_tbl.json 2> gpurun_out/bench_tbl.err; echo "bench exit=$?"; cat gpurun_out/bench_tbl.err | head -80; python -c "
import json; d=json.load(open('gpurun_out/bench_tbl.json')); print(d['value'], d['ms_per_step'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])"
timeout 300 python tools/perf_experiments.py 0 > gpurun_out/perf_exp6.log 2>&1; grep -E "BENCH wgrad" gpurun_out/perf_exp6.log
