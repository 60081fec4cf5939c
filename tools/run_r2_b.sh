#!/bin/bash
# round-2 GPU pass b: FLUX parity with two peers, optimizer/dropout/eval/wavelet tests, debug-lib groups, bench + eager, profile
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_flux.py -q -m gpu -s > gpurun_out/r02_pytest_new.log 2>&1
echo "pytest(new) exit=$?" >> gpurun_out/r02_pytest_new.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_new.log | tail -12
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu > gpurun_out/r02_pytest_kernels.log 2>&1
echo "pytest(kernels) exit=$?" >> gpurun_out/r02_pytest_kernels.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_kernels.log | tail -6
VQB_KERNEL_TABLE=1 timeout 1200 python bench.py > gpurun_out/r02_bench_b.json 2> gpurun_out/r02_bench_b.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_b.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_eager_b200')}, d['e2e']['value'], d.get('eager_b200'))
PY
timeout 900 python bench.py --config gan --no-cpu-baseline > gpurun_out/r02_bench_gan_b.json 2> gpurun_out/r02_bench_gan_b.err; echo "bench gan exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_gan_b.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_eager_b200')}, d['e2e']['value'], d.get('eager_b200'))
PY
VQB_PROFILE=1 VQB_PROFILE_ROWS=45 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32.txt 2>&1
grep -E "STEP|GPU span" gpurun_out/r02_step_profile_b32.txt
