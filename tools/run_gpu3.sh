#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests -q -m gpu -s > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|patchd|cos |FAIL|Error" gpurun_out/pytest_gpu.log | tail -40
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?" >> gpurun_out/smoke.log; tail -3 gpurun_out/smoke.log
timeout 900 python bench.py --steps 5 --warmup 3 > gpurun_out/bench_b32.json 2> gpurun_out/bench_b32.err; echo "bench exit=$?"; cat gpurun_out/bench_b32.json; tail -5 gpurun_out/bench_b32.err
timeout 600 python tools/step_bench.py 16 128 > gpurun_out/step_b16.log 2>&1; tail -2 gpurun_out/step_b16.log
timeout 600 python tools/step_bench.py 8 128 gan > gpurun_out/step_b8_gan.log 2>&1; tail -2 gpurun_out/step_b8_gan.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 2100 -c 750 --csv --log-file gpurun_out/launches_b8.csv python bench.py --steps 1 --warmup 3 --batch 8 --no-cpu-baseline > gpurun_out/ncu_bench.log 2>&1; echo "ncu list exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv_gemm_kernel -s 3 -c 1 -o gpurun_out/prof_conv_512 python tools/gpu_probe.py bench > gpurun_out/ncu_conv.log 2>&1; echo "ncu conv exit=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:wgrad_gemm_kernel -s 2 -c 1 -o gpurun_out/prof_wgrad_512 python tools/gpu_probe.py bench > gpurun_out/ncu_wgrad.log 2>&1; echo "ncu wgrad exit=$?"
ls -la gpurun_out
