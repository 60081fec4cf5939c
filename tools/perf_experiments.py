"""Perf experiments via vqb_set_debug_mode bits: &3: 1 = no TMA, 2 = no MMA; 4 = rotation off; 8 = wgrad 128-pixel K blocks;
16 = wgrad 5-D TMA maps (one request per operand)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import torch
import gpu_probe as P
L = P.L
import ctypes
L.vqb_set_debug_mode.argtypes = [ctypes.c_int]
shapes = [(8, 64, 64, 512, 512, 4), (8, 256, 256, 128, 128, 64), (8, 128, 128, 256, 256, 8), (8, 32, 32, 512, 512, 2),
          (8, 256, 256, 64, 64, 64), (8, 128, 128, 512, 256, 8)]
modes = [int(a) for a in sys.argv[1:]] or [0, 8, 16, 24, 18, 26]
for mode in modes:
    L.vqb_set_debug_mode(mode)
    print(f"##### debug mode {mode}", flush=True)
    try:
        for (N, H, W, Ci, Co, ks) in shapes:
            P.bench_wgrad(N, H, W, Ci, Co, 3, ks, iters=10)
        if (mode & 3) == 0:
            print("parity:", P.group_wgrad(), flush=True)
    except Exception as e:
        print("MODE FAILED:", repr(e)[:300], flush=True)
L.vqb_set_debug_mode(0)
