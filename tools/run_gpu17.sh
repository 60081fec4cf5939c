#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py elem 2>&1 | grep -E "FAIL|colsum|ALL PASS|HAS FAIL|Error" | head -20
VQB_KERNEL_TABLE=2 timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/bench_b32_t2.json 2> gpurun_out/bench_tbl2.err; echo "bench exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/bench_b32_t2.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 900 ncu --set full --clock-control none --profile-from-start off -k regex:'conv_gemm|wgrad_gemm|gn_bwd_reduce|gn_bwd_apply|gn_apply|colsum' -f -o /tmp/r01_resblock python tools/resblock_profile.py 128 256 16 > gpurun_out/ncu_resblock.log 2>&1; echo "ncu exit=$?"; tail -2 gpurun_out/ncu_resblock.log
ncu -i /tmp/r01_resblock.ncu-rep --page raw --csv > gpurun_out/r01_resblock_128_256_raw.csv 2>/dev/null
ls -la /tmp/r01_resblock.ncu-rep gpurun_out/
sz=$(stat -c %s /tmp/r01_resblock.ncu-rep); if [ "$sz" -lt 40000000 ]; then cp /tmp/r01_resblock.ncu-rep gpurun_out/; fi
