#!/bin/bash
mkdir -p gpurun_out
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
VQB_PROFILE=1 VQB_PROFILE_ROWS=25 timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/step_bench.py 32 128 > gpurun_out/step_profile_n2.txt 2>&1
grep -E "STEP|nccl|GPU span|idle before" gpurun_out/step_profile_n2.txt | cut -c1-70,100-210 | head -30
