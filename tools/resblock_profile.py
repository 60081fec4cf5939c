"""Two chained ResnetBlocks (fwd + bwd) at batch N: the unit that dominates the training step. Used under ncu
(--profile-from-start off; the profiled region is one forward+backward after warm-up) and for quick timing.
usage: resblock_profile.py [C] [H] [N]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "vqgan-training_b200"))
os.environ.setdefault("VQB_OFFLINE", "1")
import torch

import ae

C = int(sys.argv[1]) if len(sys.argv) > 1 else 128
H = int(sys.argv[2]) if len(sys.argv) > 2 else 256
N = int(sys.argv[3]) if len(sys.argv) > 3 else 32
torch.manual_seed(0)
m = torch.nn.Sequential(ae.ResnetBlock(C, C), ae.ResnetBlock(C, C)).cuda()
for b in m:
    torch.nn.init.normal_(b.conv2.weight, std=0.02)
x = torch.randn(N, C, H, H, device="cuda", requires_grad=True)
gy = torch.randn(N, C, H, H, device="cuda")


def step():
    y = m(x)
    y.backward(gy)


for _ in range(2):
    step()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
torch.cuda.profiler.start()
e0.record()
step()
e1.record()
torch.cuda.synchronize()
torch.cuda.profiler.stop()
print(f"resblock x2 C={C} {H}x{H} N={N}: {e0.elapsed_time(e1):.2f} ms fwd+bwd")
