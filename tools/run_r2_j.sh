#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py tests/test_gpu_train.py -q -m gpu -k "conv or stats or fat or up or large_mean or fused_groupnorm" 2>&1 | tail -3
for v in 1 0 1 0; do
VQB_LEAN_ISSUE=$v timeout 900 python bench.py --no-eager --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('lean=$v:', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']), round(d['roofline']['ms_per_step'],2))"
done
python tools/gpu_probe.py halobench 2>&1 | grep -E "dbg=0\] BENCH" | head -12
