#!/bin/bash
# round-end validation on one GPU: parity tests, smoke, contract bench (with CPU baseline), launch list, ncu captures
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py > gpurun_out/r01_bench_b32_final.json 2> gpurun_out/bench_final.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r01_bench_b32_final.json')); print(d['value'], d['ms_per_step'], d['e2e'], d['clocks'], d['gpu_launches'], d['cpu_baseline'])"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file gpurun_out/r01_launches_b32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu launches exit=$?"
wc -l gpurun_out/r01_launches_b32.csv
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'conv_gemm|wgrad_gemm|gn_bwd_reduce|gn_bwd_apply|gn_apply' -f -o /tmp/rb256 python tools/resblock_profile.py 256 128 16 > gpurun_out/ncu_resblock256.log 2>&1; echo "ncu256 exit=$?"
ncu -i /tmp/rb256.ncu-rep --page raw --csv > gpurun_out/r01_ncu_resblock_256_128_raw.csv 2>/dev/null
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'conv_gemm|wgrad_gemm' -f -o /tmp/rb128 python tools/resblock_profile.py 128 256 16 > gpurun_out/ncu_resblock128.log 2>&1; echo "ncu128 exit=$?"
ncu -i /tmp/rb128.ncu-rep --page raw --csv > gpurun_out/r01_ncu_resblock_128_256_v2_raw.csv 2>/dev/null
ls -la gpurun_out/*.csv
