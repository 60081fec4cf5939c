#!/bin/bash
mkdir -p gpurun_out
VQB_KERNEL_TABLE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32_t2.json 2> gpurun_out/bench_tbl2.err; echo "bench exit=$?"
timeout 300 python tools/resblock_profile.py 128 256 32
timeout 900 ncu --set full --import-source on --clock-control none --profile-from-start off -f -o gpurun_out/r01_resblock_128_256 python tools/resblock_profile.py 128 256 32 > gpurun_out/ncu_resblock.log 2>&1; echo "ncu exit=$?"; tail -3 gpurun_out/ncu_resblock.log
ls -la gpurun_out/*.ncu-rep
