#!/bin/bash
mkdir -p gpurun_out
VQB_DEBUG_LIB=1 timeout 300 python tools/ncu_conv128.py 32 > gpurun_out/r02_conv128_modes_after.txt 2>&1; cat gpurun_out/r02_conv128_modes_after.txt
VQB_DEBUG_LIB=1 timeout 300 python tools/ncu_conv128.py 32 256 128 >> gpurun_out/r02_conv128_modes_after.txt 2>&1; tail -3 gpurun_out/r02_conv128_modes_after.txt
VQB_DEBUG_LIB=1 timeout 600 python tools/gpu_probe.py halobench > gpurun_out/r02_halobench_after.txt 2>&1; grep -E "BENCH|dbg" gpurun_out/r02_halobench_after.txt | head -40
VQB_DEBUG_LIB=1 timeout 600 python tools/gpu_probe.py wgbench > gpurun_out/r02_wgbench_after.txt 2>&1; grep -E "BENCH|dbg|--" gpurun_out/r02_wgbench_after.txt | head -50
VQB_KERNEL_TABLE=1 timeout 900 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r02_bench_f.json 2> gpurun_out/r02_bench_f.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_f.json'))
print({k:d[k] for k in ('value','ms_per_step')}, d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])
PY
grep -E "^\('conv'|^\('wgrad'" gpurun_out/r02_bench_f.err | head -24
