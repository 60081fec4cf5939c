#!/bin/bash
mkdir -p gpurun_out
for ov in 1 0; do
VQB_DDP_OVERLAP=$ov timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 2951$ov tools/step_bench.py 32 128 2>&1 | grep STEP | sed "s/^/overlap=$ov /"
done
VQB_DDP_BUCKET_MB=25 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/step_bench.py 32 128 2>&1 | grep STEP | sed "s/^/overlap=1 bucket=25 /"
