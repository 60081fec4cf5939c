#!/bin/bash
mkdir -p gpurun_out
VQB_ISSUE2=1 timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv or stats or up" 2>&1 | tail -3
for v in 0 1 0 1; do
VQB_ISSUE2=$v VQB_KERNEL_TABLE=1 timeout 900 python bench.py --no-eager --no-cpu-baseline --steps 10 2>gpurun_out/r02_bench_k_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('issue2=$v:', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['roofline']['ms_per_step'],2))"
done
grep -E "^\('conv', 32, 256, 256, 128, 128, 9" gpurun_out/r02_bench_k_0.err gpurun_out/r02_bench_k_1.err
VQB_ISSUE2=1 python tools/gpu_probe.py halobench 2>&1 | grep -E "dbg=0\] BENCH" | head -2
VQB_ISSUE2=0 python tools/gpu_probe.py halobench 2>&1 | grep -E "dbg=0\] BENCH" | head -2
