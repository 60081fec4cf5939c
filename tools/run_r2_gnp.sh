#!/bin/bash
mkdir -p gpurun_out
O=gpurun_out/r02_gn_bwd_variants.txt
: > $O
VQB_GN_BWD_PERSISTENT=0 python tools/gn_bwd_bench.py >> $O 2>&1
for d in 0 1; do for h in 0 1; do
VQB_GN_BWD_PERSISTENT=1 VQB_GNP_DEPTH=$d VQB_GNP_HINTS=$h VQB_GNP_MB=36 python tools/gn_bwd_bench.py >> $O 2>&1
done; done
VQB_GN_BWD_PERSISTENT=1 VQB_GNP_DEPTH=0 VQB_GNP_HINTS=0 VQB_GNP_MB=18 python tools/gn_bwd_bench.py >> $O 2>&1
VQB_GN_BWD_PERSISTENT=1 VQB_GNP_DEPTH=1 VQB_GNP_HINTS=1 VQB_GNP_MB=18 python tools/gn_bwd_bench.py >> $O 2>&1
VQB_GN_BWD_PERSISTENT=1 VQB_GNP_DEPTH=0 VQB_GNP_HINTS=1 VQB_GNP_MB=72 python tools/gn_bwd_bench.py >> $O 2>&1
grep -E "total|256x256 C=128" $O
timeout 600 python -m pytest tests/test_gpu_kernels.py -q -m gpu -k "elem" 2>&1 | tail -2
