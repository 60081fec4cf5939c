#!/bin/bash
# Runs the bring-up probe groups, each in its own process under a timeout (a trapped kernel kills the context).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv > gpurun_out/probe_gpu.txt 2>&1
for g in "$@"; do
  echo "##### group $g" | tee -a gpurun_out/probe.log
  timeout 300 python tools/gpu_probe.py $g >> gpurun_out/probe.log 2>&1
  echo "##### group $g exit=$?" | tee -a gpurun_out/probe.log
done
tail -n 150 gpurun_out/probe.log
