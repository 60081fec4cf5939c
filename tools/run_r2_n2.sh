#!/bin/bash
# 2-GPU pass: NCCL gradient-equivalence test, contract bench at N=2 (CUDA graph incl. NCCL / eager launch / overlapped ranges)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name --format=csv,noheader | head -4
timeout 900 python -m pytest tests/test_gpu_train.py -q -m gpu -s -k "two_rank" > gpurun_out/r02_pytest_n2.log 2>&1
echo "pytest(n2) exit=$?" >> gpurun_out/r02_pytest_n2.log
grep -E "passed|failed|skipped|N=2 vs|exit=" gpurun_out/r02_pytest_n2.log | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-eager > gpurun_out/r02_scale_n2_graph.json 2> gpurun_out/r02_scale_n2_graph.err; echo "n2 graph exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n2_graph.json')); print('N=2 graph:', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])" || tail -20 gpurun_out/r02_scale_n2_graph.err
timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-eager --no-graph 2>gpurun_out/r02_scale_n2_nograph.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 no graph:', d['value'], d['ms_per_step'], d['e2e']['value'])"
VQB_DDP_OVERLAP=2 timeout 900 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-eager 2>gpurun_out/r02_scale_n2_overlap.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 graph + 2-range overlap:', d['value'], d['ms_per_step'], d['e2e']['value'])"
timeout 600 python bench.py --no-eager --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box:', d['value'], d['ms_per_step'])"
timeout 900 $TR bench.py --gpus 2 --steps 6 --warmup 3 2>gpurun_out/r02_scale_n2_eager.err | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=2 with eager peer:', d['value'], d.get('eager_b200'), d.get('vs_eager_b200'))"
