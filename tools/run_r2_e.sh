#!/bin/bash
mkdir -p gpurun_out
python tools/gn_bwd_bench.py > gpurun_out/r02_gn_bwd_trimmed.txt 2>&1; grep -E "total" gpurun_out/r02_gn_bwd_trimmed.txt
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/r02_pytest_kernels_e.log 2>&1
echo "pytest(kernels) exit=$?" >> gpurun_out/r02_pytest_kernels_e.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_kernels_e.log | tail -6
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_flux.py tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/r02_pytest_e.log 2>&1
echo "pytest exit=$?" >> gpurun_out/r02_pytest_e.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_e.log | tail -12
timeout 900 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r02_bench_e.json 2> gpurun_out/r02_bench_e.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_e.json'))
print('graph:', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config']['cuda_graph'])
PY
VQB_CUDA_GRAPH=0 VQB_PROFILE=1 VQB_PROFILE_ROWS=40 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32_e.txt 2>&1
grep -E "STEP|GPU span|gn_bwd|pack_weights" gpurun_out/r02_step_profile_b32_e.txt | cut -c1-70,150-235
