#!/bin/bash
# final evidence refresh with the committed code: gpu tests, smoke, contract bench (all legs)
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r02_pytest_gpu_final.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_gpu_final.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
VQB_KERNEL_TABLE=1 timeout 900 python bench.py --gpus 1 --steps 10 --warmup 3 > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','steps','warmup','gpu_launches','vs_eager_b200')}, d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline']['frac'], d['roofline_wgrad']['achieved'], d['eager_b200'].get('value'), d['cpu_baseline']['value'])
PY
timeout 600 python bench.py --impl reference --gpus 1 --steps 3 --warmup 1 | cut -c1-400
VQB_CUDA_GRAPH=0 VQB_PROFILE=1 VQB_PROFILE_ROWS=45 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32_final.txt 2>&1
grep -E "STEP|GPU span" gpurun_out/r02_step_profile_b32_final.txt
