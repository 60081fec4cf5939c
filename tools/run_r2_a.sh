#!/bin/bash
# round-2 first GPU pass: new FLUX-config parity tests + optimizer/repack tests, full gpu suite, smoke, contract bench with eager peer
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,memory.total --format=csv,noheader > gpurun_out/r02_gpus.txt
timeout 1500 python -m pytest tests/test_gpu_train.py tests/test_gpu_flux.py -q -m gpu -s > gpurun_out/r02_pytest_new.log 2>&1
echo "pytest(new) exit=$?" >> gpurun_out/r02_pytest_new.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_new.log | tail -12
timeout 900 python -m pytest tests -q -m gpu --deselect tests/test_gpu_flux.py --deselect tests/test_gpu_train.py > gpurun_out/r02_pytest_old.log 2>&1
echo "pytest(old) exit=$?" >> gpurun_out/r02_pytest_old.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/r02_pytest_old.log | tail -6
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 1200 python bench.py > gpurun_out/r02_bench_a.json 2> gpurun_out/r02_bench_a.err; echo "bench exit=$?"
tail -c 3000 gpurun_out/r02_bench_a.json
tail -5 gpurun_out/r02_bench_a.err
