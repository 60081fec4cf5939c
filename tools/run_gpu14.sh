#!/bin/bash
mkdir -p gpurun_out
VQB_KERNEL_TABLE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32_t2.json 2> gpurun_out/bench_tbl2.err; echo "bench exit=$?"
