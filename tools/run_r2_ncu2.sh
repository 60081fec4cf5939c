#!/bin/bash
mkdir -p gpurun_out
timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_gemm|wgrad_gemm' -f -o gpurun_out/r02_conv_wgrad_after python tools/ncu_conv_wgrad.py 16 > gpurun_out/r02_ncu2.log 2>&1; echo "ncu exit=$?"
ncu -i gpurun_out/r02_conv_wgrad_after.ncu-rep --page raw --csv > gpurun_out/r02_ncu_conv_wgrad_after_raw.csv 2>/dev/null
ncu -i gpurun_out/r02_conv_wgrad_after.ncu-rep --page source --csv --print-source sass > gpurun_out/r02_ncu_conv_wgrad_after_source.csv 2>/dev/null
ls -la gpurun_out/r02_conv_wgrad_after.ncu-rep gpurun_out/r02_ncu_conv_wgrad_after_raw.csv gpurun_out/r02_ncu_conv_wgrad_after_source.csv
