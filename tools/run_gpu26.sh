#!/bin/bash
mkdir -p gpurun_out
timeout 300 python tools/gpu_probe.py tinybench 2>&1 | grep -E "BENCH|Error|error|Traceback"
timeout 600 ncu --set full --clock-control none -k regex:'conv_gemm|wgrad_gemm' -c 12 -f -o /tmp/tiny python tools/gpu_probe.py tinybench > gpurun_out/ncu_tiny.log 2>&1; echo "ncu exit=$?"
ncu -i /tmp/tiny.ncu-rep --page raw --csv > gpurun_out/r01_ncu_tiny_raw.csv 2>/dev/null; ls -la gpurun_out/r01_ncu_tiny_raw.csv
