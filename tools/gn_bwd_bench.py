"""GroupNorm(+swish) backward micro-benchmark at the step's shapes (N=32): time and effective bytes/element.
Variants through env: VQB_GN_BWD_PERSISTENT (0 two kernels / 1 persistent), VQB_GNP_DEPTH, VQB_GNP_HINTS, VQB_GNP_MB."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vqgan-training_b200"))
os.environ.setdefault("VQB_OFFLINE", "1")
import torch

import ops

N = int(sys.argv[1]) if len(sys.argv) > 1 else 32
tag = " ".join(f"{k}={os.environ[k]}" for k in ("VQB_GN_BWD_PERSISTENT", "VQB_GNP_DEPTH", "VQB_GNP_HINTS", "VQB_GNP_MB")
               if k in os.environ)
tot = 0.0
for (H, C, add) in [(256, 128, False), (256, 128, True), (128, 256, False), (128, 256, True), (64, 512, False),
                    (64, 512, True), (32, 512, True), (256, 256, True)]:
    torch.manual_seed(0)
    x = (torch.randn(N, H, H, C, device="cuda") * 1.3 + 0.2).to(torch.bfloat16).requires_grad_(True)
    ga = (1 + 0.1 * torch.randn(C, device="cuda")).requires_grad_(True)
    be = (0.1 * torch.randn(C, device="cuda")).requires_grad_(True)
    gy = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)
    gs = torch.randn(N, H, H, C, device="cuda").to(torch.bfloat16)

    def run():
        if add:
            y, sk = ops.group_norm_silu(x, ga, be, 32, 1e-6, True, with_skip=True)
            return torch.autograd.grad([y, sk], [x, ga, be], [gy, gs])
        y = ops.group_norm_silu(x, ga, be, 32, 1e-6, True)
        return torch.autograd.grad([y], [x, ga, be], [gy])

    # time the backward only: build the graph outside the timed region
    ts = []
    for it in range(6):
        if add:
            y, sk = ops.group_norm_silu(x, ga, be, 32, 1e-6, True, with_skip=True)
            outs, grads = [y, sk], [gy, gs]
        else:
            y = ops.group_norm_silu(x, ga, be, 32, 1e-6, True)
            outs, grads = [y], [gy]
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        g = torch.autograd.grad(outs, [x, ga, be], grads)
        e1.record()
        torch.cuda.synchronize()
        if it >= 2:
            ts.append(e0.elapsed_time(e1))
    ms = sum(ts) / len(ts)
    el = N * H * H * C
    tot += ms
    print(f"[{tag}] gn_bwd {H}x{H} C={C} add={int(add)}: {ms * 1e3:8.1f} us  = {el * (8 if add else 6) / ms / 1e9:6.2f} TB/s at the "
          f"{8 if add else 6} B/elem minimum", flush=True)
print(f"[{tag}] total {tot:.3f} ms")
