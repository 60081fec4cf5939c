#!/bin/bash
mkdir -p gpurun_out
VQB_PROFILE=1 timeout 600 python tools/step_bench.py 32 128 > gpurun_out/step_profile_b32.txt 2>&1
tail -75 gpurun_out/step_profile_b32.txt | cut -c1-200
