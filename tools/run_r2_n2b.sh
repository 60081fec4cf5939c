#!/bin/bash
# 2-GPU validation: NCCL gradient-equivalence test; contract bench exactly as the driver launches it (eager peer leg
# included); opt-in multi-rank CUDA graph incl. clean teardown. Short timeouts: a hang must not eat the budget.
mkdir -p gpurun_out
timeout 400 python -m pytest tests/test_gpu_train.py -q -m gpu -s -k "two_rank" > gpurun_out/r02_pytest_n2.log 2>&1
echo "pytest(n2) exit=$?" >> gpurun_out/r02_pytest_n2.log
grep -E "passed|failed|skipped|N=2 vs|exit=" gpurun_out/r02_pytest_n2.log | tail -5
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511"
timeout 420 $TR bench.py --gpus 2 --steps 10 --warmup 3 > gpurun_out/r02_scale_n2.json 2> gpurun_out/r02_scale_n2.err; echo "n2 default exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n2.json')); print('N=2 default:', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'], d.get('eager_b200'), d.get('vs_eager_b200'))" || tail -20 gpurun_out/r02_scale_n2.err
VQB_CUDA_GRAPH=2 timeout 300 $TR bench.py --gpus 2 --steps 10 --warmup 3 --no-eager > gpurun_out/r02_scale_n2_graph.json 2> gpurun_out/r02_scale_n2_graph.err; echo "n2 graph exit=$?"
python -c "
import json; d=json.load(open('gpurun_out/r02_scale_n2_graph.json')); print('N=2 graph:', d['value'], d['ms_per_step'], d['e2e']['value'], d['config']['cuda_graph'])"
timeout 300 python bench.py --no-eager --no-cpu-baseline --steps 10 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('N=1 same box:', d['value'], d['ms_per_step'])"
timeout 300 $TR bench.py --impl reference --gpus 2 --steps 2 --warmup 1 2>/dev/null | cut -c1-300
