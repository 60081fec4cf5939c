#!/bin/bash
mkdir -p gpurun_out
VQB_KERNEL_TABLE=1 timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32.json 2> gpurun_out/bench_tbl.err; echo "bench exit=$?"; grep -E ", 8, |, 3, 9|, 24, " gpurun_out/bench_tbl.err | head; python -c "
import json; d=json.load(open('gpurun_out/bench_b32.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])"
VQB_FAT_CONV=0 timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -3
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -3
