#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/perf_experiments.py 0 > gpurun_out/perf_exp9.log 2>&1; grep -E "BENCH|parity|rror" gpurun_out/perf_exp9.log | cut -c1-150
timeout 900 python -m pytest tests -q -m gpu --durations=8 > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=|s call|s setup" gpurun_out/pytest_gpu.log | tail -16
VQB_KERNEL_TABLE=1 timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_tbl.err; echo "bench exit=$?"; head -24 gpurun_out/bench_tbl.err; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'], d.get('cpu_baseline'))"
