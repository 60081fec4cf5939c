"""One profiled launch each of: conv 3x3 128->128 @ 256^2 (halo, N = 128 tiles), conv 3x3 256->256 @ 128^2 (halo, N = 256),
wgrad 128->128 @ 256^2, wgrad 512->512 @ 32^2 — for `ncu --set full --import-source on --profile-from-start off`."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vqgan-training_b200"))
os.environ.setdefault("VQB_OFFLINE", "1")
import torch

import ops
import plans

N = int(sys.argv[1]) if len(sys.argv) > 1 else 16
torch.manual_seed(0)
for (C, H) in [(128, 256), (256, 128)]:
    x = (torch.randn(N, H, H, C, device="cuda") * 0.5).to(torch.bfloat16)
    w = torch.randn(C, C, 3, 3, device="cuda") * 0.03
    b = torch.randn(C, device="cuda") * 0.1
    cache = ops.PackedCache()
    with torch.no_grad():
        for _ in range(3):
            y = ops.conv(x, w, b, cache, "s1")
        torch.cuda.synchronize()
        torch.cuda.profiler.start()
        y = ops.conv(x, w, b, cache, "s1")
        torch.cuda.synchronize()
        torch.cuda.profiler.stop()
for (C, H) in [(128, 256), (512, 32)]:
    x = (torch.randn(N, H, H, C, device="cuda") * 0.5).to(torch.bfloat16)
    dy = (torch.randn(N, H, H, C, device="cuda") * 0.5).to(torch.bfloat16)
    g = plans.geom_s1(N, H, H, C, 3)
    for _ in range(3):
        gw = ops.run_wgrad(g, x, dy, (C, C, 3, 3), C)
    torch.cuda.synchronize()
    torch.cuda.profiler.start()
    gw = ops.run_wgrad(g, x, dy, (C, C, 3, 3), C)
    torch.cuda.synchronize()
    torch.cuda.profiler.stop()
print("done")
