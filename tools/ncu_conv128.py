"""One profiled launch per mode of the 3x3 128->128 @ 256^2 conv (forward, bias epilogue): default halo mode (two
independent N = 128 accumulators per CTA), CTA-pair mode (cta_group::2, debug bit 8192) and swap mode (weights as M,
256 pixels as N, debug bit 4096). Needs the -DVQB_DEBUG library: VQB_DEBUG_LIB=1. Used under
`ncu --set full --profile-from-start off`. usage: ncu_conv128.py [N] [C] [H]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vqgan-training_b200"))
os.environ.setdefault("VQB_OFFLINE", "1")
os.environ["VQB_DEBUG_LIB"] = "1"
import torch

import native
import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
C = int(sys.argv[2]) if len(sys.argv) > 2 else 128
H = int(sys.argv[3]) if len(sys.argv) > 3 else 256
L = native.load()
torch.manual_seed(0)
x = (torch.randn(N, H, H, C, device="cuda") * 0.5).to(torch.bfloat16)
w = torch.randn(C, C, 3, 3, device="cuda") * 0.03
b = torch.randn(C, device="cuda") * 0.1
for mode in (0, 8192, 4096):
    assert L.vqb_set_debug_mode(mode) == 0
    cache = ops.PackedCache()
    with torch.no_grad():
        for _ in range(3):
            y = ops.conv(x, w, b, cache, "s1")
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(5):
            y = ops.conv(x, w, b, cache, "s1")
        e1.record()
        torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        print(f"mode {mode}: {ms * 1e3:.1f} us = {2.0 * N * H * H * C * C * 9 / ms / 1e9:.1f} TFLOP/s", flush=True)
        torch.cuda.profiler.start()
        y = ops.conv(x, w, b, cache, "s1")
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
L.vqb_set_debug_mode(0)
