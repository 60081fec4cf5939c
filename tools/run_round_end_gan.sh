#!/bin/bash
mkdir -p gpurun_out
timeout 600 python bench.py --gan --batch 32 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r01_bench_gan_b32_final.json 2> gpurun_out/bench_gan.err; echo "bench gan exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r01_bench_gan_b32_final.json')); print(d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'])"
tail -3 gpurun_out/bench_gan.err | cut -c1-200
