"""Perf experiments via vqb_set_debug_mode bits: &3: 1 = no TMA, 2 = no MMA; 4 = rotation off; 8 = wgrad 64-pixel K blocks;
16 = wgrad 4-D maps (one TMA per atom); 32 = conv: force 128-pixel tiles (no double-M tiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gpu_probe as P
L = P.L
import ctypes
L.vqb_set_debug_mode.argtypes = [ctypes.c_int]
shapes = [(8, 64, 64, 512, 512), (8, 256, 256, 128, 128), (8, 128, 128, 256, 256), (8, 32, 32, 512, 512),
          (8, 256, 256, 64, 64), (8, 128, 128, 512, 256), (32, 32, 32, 512, 512), (32, 64, 64, 512, 512), (8, 256, 256, 256, 256)]
modes = [int(a) for a in sys.argv[1:]] or [32, 0]
for mode in modes:
    L.vqb_set_debug_mode(mode)
    print(f"##### debug mode {mode}", flush=True)
    try:
        for (N, H, W, Ci, Co) in shapes:
            P.bench_conv(N, H, W, Ci, Co, 3, iters=10)
        P.bench_conv(8, 128, 128, 512, 256, 1, iters=10)
        P.bench_conv(8, 256, 256, 256, 128, 1, iters=10)
        if (mode & 3) == 0:
            print("parity:", P.group_gemm() and P.group_conv() and P.group_conv2(), flush=True)
    except Exception as e:
        print("MODE FAILED:", repr(e)[:300], flush=True)
L.vqb_set_debug_mode(0)
P.bench_wgrad(8, 64, 64, 512, 512, 3, 4)
P.bench_wgrad(8, 256, 256, 128, 128, 3, 64)
