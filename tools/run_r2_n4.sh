#!/bin/bash
# N-GPU validation exactly as the driver launches the contract bench (CUDA graph incl. NCCL, eager peer leg, clean teardown)
N=${1:-4}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
timeout 420 $TR bench.py --gpus $N --steps 10 --warmup 3 > gpurun_out/r02_scale_n$N.json 2> gpurun_out/r02_scale_n$N.err; echo "n$N exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n$N.json')); print('N=$N:', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d.get('eager_b200',{}).get('value'), d.get('vs_eager_b200'))" || tail -30 gpurun_out/r02_scale_n$N.err
timeout 300 python bench.py --no-eager --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box:', d['value'], d['ms_per_step'])"
