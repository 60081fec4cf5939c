#!/bin/bash
# compute-sanitizer memcheck (all kernel parity groups) + racecheck (shared-memory hazards) logs -> profiles/
mkdir -p gpurun_out
OUT=gpurun_out/r02_sanitizer.txt
echo "# compute-sanitizer runs of tools/gpu_probe.py groups (product libvqb200.so), $(date -u +%FT%TZ)" > $OUT
for g in gemm conv2 wgrad elem lpips up stats fat; do
  echo "== memcheck group $g" >> $OUT
  timeout 600 compute-sanitizer --tool memcheck --error-exitcode 9 --print-limit 5 python tools/gpu_probe.py $g > gpurun_out/san_$g.log 2>&1
  echo "exit=$?" >> $OUT
  grep -E "ERROR SUMMARY|== group|Invalid|out of bounds|misaligned" gpurun_out/san_$g.log | tail -5 >> $OUT
done
for g in gemm elem stats; do
  echo "== racecheck group $g" >> $OUT
  timeout 900 compute-sanitizer --tool racecheck --error-exitcode 9 --print-limit 5 python tools/gpu_probe.py $g > gpurun_out/race_$g.log 2>&1
  echo "exit=$?" >> $OUT
  grep -E "RACECHECK SUMMARY|== group|hazard" gpurun_out/race_$g.log | tail -5 >> $OUT
done
cat $OUT
