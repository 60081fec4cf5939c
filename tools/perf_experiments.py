"""Perf experiments via vqb_set_debug_mode bits: &3: 1 = no TMA, 2 = no MMA; 4 = rotation off; 8 = wgrad 64-pixel K blocks;
16 = wgrad 4-D maps; 32 = conv: force 128-pixel tiles; 64 = wgrad: force 128-row tiles; 128 = conv: skip epilogue stores;
256 = conv: direct (non-TMA) epilogue stores."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gpu_probe as P
L = P.L
shapes = [(32, 128, 128, 128, 256, 1), (32, 256, 256, 256, 128, 1), (32, 256, 256, 128, 128, 3), (32, 256, 256, 8, 128, 3),
          (32, 32, 32, 512, 512, 3), (32, 64, 64, 512, 512, 3), (32, 128, 128, 256, 256, 3), (32, 256, 256, 64, 64, 3)]
modes = [int(a) for a in sys.argv[1:]] or [0, 256]
print("parity (TMA-store epilogue):", P.group_gemm() and P.group_conv() and P.group_conv2(), flush=True)
for mode in modes:
    L.vqb_set_debug_mode(mode)
    print(f"##### debug mode {mode}", flush=True)
    for (N, H, W, Ci, Co, k) in shapes:
        P.bench_conv(N, H, W, Ci, Co, k, iters=5)
L.vqb_set_debug_mode(0)
