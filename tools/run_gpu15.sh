#!/bin/bash
mkdir -p gpurun_out
cd vqgan-training_b200 2>/dev/null; cd ..
timeout 300 python tools/gpu_probe.py conv 2>&1 | tail -22
timeout 300 python tools/gpu_probe.py conv2 2>&1 | tail -4
timeout 300 python tools/gpu_probe.py resbench 2>&1 | grep -E "BENCH|Error|error" 
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
