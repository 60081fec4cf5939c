#!/bin/bash
mkdir -p gpurun_out
VQB_HALO_MIN_COUT=64 timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "conv or stats or fat" 2>&1 | tail -3
for v in 128 64 128 64; do
VQB_HALO_MIN_COUT=$v VQB_KERNEL_TABLE=1 timeout 900 python bench.py --no-eager --no-cpu-baseline --steps 10 2> gpurun_out/r02_bench_g_$v.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('halo_min_cout=$v:', round(d['value'],1), round(d['ms_per_step'],2), d['clocks']['sm_mhz'], round(d['roofline']['achieved']))"
done
grep -E "^\('conv', 32, 256, 256, 64, 64, 9|^\('conv', 32, 128, 128, 64, 128, 9|^\('conv', 32, 128, 128, 128, 64, 9" gpurun_out/r02_bench_g_128.err gpurun_out/r02_bench_g_64.err
VQB_HALO_MIN_COUT=64 timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_flux.py -q -m gpu 2>&1 | tail -3
