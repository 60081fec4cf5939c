#!/bin/bash
# weak-scaling check on 2 GPUs with the contract bench (one rank per GPU over NCCL); stdout must be exactly one JSON line
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 2 --steps 6 --warmup 3 --no-cpu-baseline > gpurun_out/r01_scale_n2_b32_final.json 2> gpurun_out/bench_n2.err; echo "bench n2 exit=$?"
wc -l gpurun_out/r01_scale_n2_b32_final.json
python -c "
import json; d=json.load(open('gpurun_out/r01_scale_n2_b32_final.json')); print(d['n_gpus'], d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'])"
