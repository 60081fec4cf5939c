#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt
timeout 600 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit=$?" >> gpurun_out/pytest_gpu.log
grep -E "passed|failed|FAIL|Error|exit=" gpurun_out/pytest_gpu.log | tail -6
timeout 600 python bench.py --gpus 1 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/scale_n1.json 2> gpurun_out/scale_n1.err; echo "n1 exit=$?"
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/scale_n2.json 2> gpurun_out/scale_n2.err; echo "n2 exit=$?"; tail -3 gpurun_out/scale_n2.err
python - <<'PY'
import json
for n in (1,2):
    try:
        d=json.loads(open(f'gpurun_out/scale_n{n}.json').read().strip().splitlines()[-1])
        print(n, d['value'], d['ms_per_step'], d['e2e']['value'], d['clocks'], d['roofline']['achieved'], d['roofline_wgrad']['achieved'])
    except Exception as e: print(n, 'ERR', e)
PY
