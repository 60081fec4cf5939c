#!/bin/bash
# GPU pass c: fused GN-backward path — kernel test, FLUX parity, bench (no eager/cpu legs), step profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -s -k "fused_groupnorm or packed or adamw" > gpurun_out/r02_pytest_fuse.log 2>&1
echo "pytest(fuse) exit=$?" >> gpurun_out/r02_pytest_fuse.log
grep -E "passed|failed|FAIL|Error|exit=|fused vs" gpurun_out/r02_pytest_fuse.log | tail -30
timeout 1500 python -m pytest tests/test_gpu_flux.py tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/r02_pytest_flux.log 2>&1
echo "pytest(flux) exit=$?" >> gpurun_out/r02_pytest_flux.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_flux.log | tail -8
VQB_KERNEL_TABLE=1 timeout 900 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r02_bench_c.json 2> gpurun_out/r02_bench_c.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_c.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['roofline']['achieved'], d['roofline']['ms_per_step'], d['roofline_wgrad']['achieved'])
PY
VQB_GN_BWD_FUSE=0 timeout 900 python bench.py --no-eager --no-cpu-baseline --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('unfused:', d['value'], d['ms_per_step'])"
VQB_PROFILE=1 VQB_PROFILE_ROWS=45 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32_fuse.txt 2>&1
grep -E "STEP|GPU span" gpurun_out/r02_step_profile_b32_fuse.txt
