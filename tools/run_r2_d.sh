#!/bin/bash
# GPU pass d: CUDA-graph step + persistent GN backward: kernel groups, train/flux tests, bench graph vs eager, profile
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "elem or stats or up or conv2" > gpurun_out/r02_pytest_kernels_d.log 2>&1
echo "pytest(kernels) exit=$?" >> gpurun_out/r02_pytest_kernels_d.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_kernels_d.log | tail -6
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_flux.py tests/test_gpu_parity.py -q -m gpu -s > gpurun_out/r02_pytest_d.log 2>&1
echo "pytest exit=$?" >> gpurun_out/r02_pytest_d.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_d.log | tail -12
timeout 900 python bench.py --no-eager --no-cpu-baseline > gpurun_out/r02_bench_d.json 2> gpurun_out/r02_bench_d.err; echo "bench exit=$?"
tail -3 gpurun_out/r02_bench_d.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_d.json'))
print('graph:', {k:d[k] for k in ('value','ms_per_step','gpu_launches')}, d['e2e']['value'], d['config']['cuda_graph'])
PY
timeout 900 python bench.py --no-eager --no-cpu-baseline --no-graph --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('eager-launch:', d['value'], d['ms_per_step'], d['e2e']['value'])"
VQB_GN_BWD_PERSISTENT=0 timeout 900 python bench.py --no-eager --no-cpu-baseline --steps 6 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('graph, 2-kernel GN bwd:', d['value'], d['ms_per_step'])"
timeout 900 python bench.py --config gan --batch 16 --no-eager --no-cpu-baseline --steps 8 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gan B=16 graph:', d['value'], d['ms_per_step'], d['e2e']['value'])"
VQB_CUDA_GRAPH=0 VQB_PROFILE=1 VQB_PROFILE_ROWS=40 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32_d.txt 2>&1
grep -E "STEP|GPU span" gpurun_out/r02_step_profile_b32_d.txt
