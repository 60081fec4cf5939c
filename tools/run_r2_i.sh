#!/bin/bash
mkdir -p gpurun_out
for u in 4 2 6 8; do VQB_GN_RED_U=$u python tools/gn_bwd_bench.py 2>&1 | grep -E "total|256x256 C=128 add=0|64x64 C=512 add=0" | sed "s/^/U=$u /"; done
bash tools/run_r2_sanitizer.sh > /dev/null 2>&1
cat gpurun_out/r02_sanitizer.txt
