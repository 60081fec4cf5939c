#!/bin/bash
# round-2 evidence pass: full gpu test suite + smoke, contract bench (all legs), every BASELINE config, launch list
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > gpurun_out/r02_pytest_gpu_final.log 2>&1; echo "pytest exit=$?" >> gpurun_out/r02_pytest_gpu_final.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_gpu_final.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
VQB_KERNEL_TABLE=1 timeout 900 python bench.py > gpurun_out/r02_bench_final.json 2> gpurun_out/r02_bench_final.err; echo "bench exit=$?"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_final.json'))
print({k:d[k] for k in ('value','ms_per_step','gpu_launches','vs_eager_b200')}, d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'], d['eager_b200'].get('value'), d['cpu_baseline']['value'])
PY
for cfg in gan vq; do
timeout 600 python bench.py --config $cfg --no-cpu-baseline > gpurun_out/r02_bench_$cfg.json 2> gpurun_out/r02_bench_$cfg.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_$cfg.json')); print('$cfg', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d.get('vs_eager_b200'), d['eager_b200'].get('value'), round(d['peak_mem_gib'],1))"
done
for b in 1 2 4 8 16; do
EX="--no-eager"; [ $b = 8 ] && EX=""
timeout 600 python bench.py --config hr512 --batch $b --no-cpu-baseline --steps 8 $EX > gpurun_out/r02_bench_hr512_b$b.json 2> gpurun_out/r02_bench_hr512_b$b.err
python -c "
import json; d=json.load(open('gpurun_out/r02_bench_hr512_b$b.json')); print('hr512 B=$b', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1), d.get('vs_eager_b200'), round(d['peak_mem_gib'],1), round(d['step_frac_of_peak'],3))"
done
timeout 600 python bench.py --config gan --batch 16 --no-cpu-baseline --no-eager --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gan B=16', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))"
timeout 600 python bench.py --config gan --batch 8 --no-cpu-baseline --no-eager --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('gan B=8', round(d['value'],1), round(d['ms_per_step'],2), round(d['e2e']['value'],1))"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches_b32.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-eager --no-graph > gpurun_out/r02_ncu_bench.log 2>&1; echo "ncu launches exit=$?"; wc -l gpurun_out/r02_launches_b32.csv
VQB_CUDA_GRAPH=0 VQB_PROFILE=1 VQB_PROFILE_ROWS=45 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/r02_step_profile_b32_final.txt 2>&1
grep -E "STEP|GPU span" gpurun_out/r02_step_profile_b32_final.txt
