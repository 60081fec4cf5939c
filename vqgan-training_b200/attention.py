"""Self-attention core of ae.AttnBlock on the flash-style warp-MMA kernels of csrc/attention.cu."""
from __future__ import annotations

import torch

import native
from native import check, ptr, stream_ptr


class MHSAFn(torch.autograd.Function):
    """qkv [N,H,W,3C] bf16 (q | k | v channel blocks) -> softmax(q k^T / sqrt(64)) v as [N,H,W,C] bf16
    (F.scaled_dot_product_attention + rearranges of ae.py:79-89)."""

    @staticmethod
    def forward(ctx, qkv, heads, head_dim):
        assert head_dim == 64, "AttnBlock.head_dim is fixed at 64 (ae.py:61)"
        qkv = qkv.contiguous()
        N, H, W, C3 = qkv.shape
        C = C3 // 3
        assert C == heads * head_dim
        out = torch.empty(N, H, W, C, device=qkv.device, dtype=torch.bfloat16)
        lse = torch.empty(N, heads, H * W, device=qkv.device, dtype=torch.float32)
        check(native.load().vqb_attn_fwd(ptr(qkv), ptr(out), ptr(lse), N, H * W, C, stream_ptr()), "attn_fwd")
        ctx.save_for_backward(qkv, out, lse)
        return out

    @staticmethod
    def backward(ctx, dout):
        qkv, out, lse = ctx.saved_tensors
        N, H, W, C3 = qkv.shape
        C = C3 // 3
        dout = dout.contiguous()
        dqkv = torch.empty_like(qkv)
        dvec = torch.empty_like(lse)
        check(native.load().vqb_attn_bwd(ptr(qkv), ptr(out), ptr(dout), ptr(lse), ptr(dvec), ptr(dqkv), N, H * W, C,
                                         stream_ptr()), "attn_bwd")
        return dqkv, None, None


def mhsa(qkv, heads, head_dim):
    return MHSAFn.apply(qkv, heads, head_dim)
