#!/bin/bash
# ncu --set full of the 128->128 @ 256^2 conv in default / pair / swap modes (debug library)
mkdir -p gpurun_out
VQB_DEBUG_LIB=1 timeout 300 python tools/ncu_conv128.py 16 > gpurun_out/r02_conv128_modes.txt 2>&1
cat gpurun_out/r02_conv128_modes.txt
VQB_DEBUG_LIB=1 timeout 1200 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:'conv_gemm' -f -o gpurun_out/r02_conv128_modes python tools/ncu_conv128.py 16 > gpurun_out/r02_ncu_conv128.log 2>&1; echo "ncu exit=$?"
ncu -i gpurun_out/r02_conv128_modes.ncu-rep --page raw --csv > gpurun_out/r02_ncu_conv128_modes_raw.csv 2>/dev/null
ls -la gpurun_out/r02_conv128_modes.ncu-rep gpurun_out/r02_ncu_conv128_modes_raw.csv
