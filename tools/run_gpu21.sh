#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py conv 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head
timeout 300 python tools/gpu_probe.py elem 2>&1 | grep -E "FAIL|ALL PASS|HAS FAIL|Error|error|Traceback" | head
timeout 600 python tools/step_bench.py 32 128 2>&1 | tail -1
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/step_profile_b32.txt 2>&1
grep -E "gn_bwd|gn_apply|conv_gemm_kernel|wgrad_gemm|colsum|wgrad_reduce|GPU span" gpurun_out/step_profile_b32.txt | cut -c1-60,100-200 | head -20
