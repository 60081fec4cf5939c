"""Geometry descriptors (views + taps) for every convolution variant on the hot path.

A "plan" is a filled VqbConvDesc / VqbWgradDesc plus the tap map used to pack the OIHW fp32 master
weights into the bf16 [rows][slot][K] matrix the tcgen05 kernels read. Plans depend only on shapes
and are cached by the modules.

Variants (reference call sites):
  s1      : k x k stride-1 "same" conv (3x3 p1, 1x1 p0)                ae.py:105-117, VGG utils.py:95-111
  s2      : 3x3 stride-2 conv after F.pad(0,1,0,1)  (Downsample)       ae.py:143-154
  patch   : k x k stride-k non-overlapping conv (PatchD heads)         utils.py:156-185
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Tuple

from native import VqbConvDesc, VqbTap, VqbView, VqbWgradDesc, dense_view


def cpad(c: int) -> int:
    """Internal NHWC channel count: multiple of 8 (16-byte pixel rows)."""
    return (c + 7) // 8 * 8


@dataclass
class ConvGeom:
    """Views/taps of the A operand for out-grid (N, Ho, Wo); tapmap[slot] = source tap in the KHxKW kernel."""
    N: int
    Ho: int
    Wo: int
    C: int  # channels of the A tensor (padded)
    views: List[VqbView] = field(default_factory=list)
    taps: List[Tuple[int, int, int]] = field(default_factory=list)  # (view, dw, dh)
    tapmap: List[int] = field(default_factory=list)
    tapmask: List[int] = field(default_factory=list)  # folded weights: slot -> bit set of source taps (else empty)


def geom_s1(N, H, W, C, k) -> ConvGeom:
    p = (k - 1) // 2
    g = ConvGeom(N, H, W, C, [dense_view(N, H, W, C)])
    for kh in range(k):
        for kw in range(k):
            g.taps.append((0, kw - p, kh - p))
            g.tapmap.append(kh * k + kw)
    return g


def geom_s1_dgrad(N, H, W, Cout_pad, k) -> ConvGeom:
    """dgrad of a stride-1 same conv = same conv over dy with rotated taps (weights packed transposed)."""
    p = (k - 1) // 2
    g = ConvGeom(N, H, W, Cout_pad, [dense_view(N, H, W, Cout_pad)])
    for kh in range(k):
        for kw in range(k):
            # slot (kh,kw) reads dy at (h + kh - p, w + kw - p) and uses source tap (k-1-kh, k-1-kw)
            g.taps.append((0, kw - p, kh - p))
            g.tapmap.append((k - 1 - kh) * k + (k - 1 - kw))
    return g


def geom_s2(N, H, W, C) -> ConvGeom:
    """3x3 stride-2 conv over x padded by one zero row/col at bottom/right: out (H/2, W/2).
    Tap (kh,kw) reads x[2ho+kh, 2wo+kw] = parity view (kh&1, kw&1) at (ho + kh//2, wo + kw//2)."""
    assert H % 2 == 0 and W % 2 == 0, "Downsample needs even H, W"
    g = ConvGeom(N, H // 2, W // 2, C)
    for ph in range(2):
        for pw in range(2):
            g.views.append(VqbView(offset=(ph * W + pw) * C, Wv=W // 2, Hv=H // 2, Nv=N, _pad=0, sw=2 * C,
                                   sh=2 * W * C, sn=H * W * C))
    for kh in range(3):
        for kw in range(3):
            g.taps.append(((kh & 1) * 2 + (kw & 1), kw // 2, kh // 2))
            g.tapmap.append(kh * 3 + kw)
    return g


def geom_s2_dgrad_classes(N, H, W, Cout_pad):
    """dgrad of the stride-2 conv, one small conv per output parity class (ph,pw) of dx (H x W):
    dx[2a+ph, 2b+pw] = sum over taps with kh%2==ph, kw%2==pw of dy[a - (kh-ph)/2, b - (kw-pw)/2] * W[kh,kw].
    Returns [(ph, pw, ConvGeom over dy grid (N, H/2, W/2))]."""
    out = []
    Ho, Wo = H // 2, W // 2
    for ph in range(2):
        for pw in range(2):
            g = ConvGeom(N, Ho, Wo, Cout_pad, [dense_view(N, Ho, Wo, Cout_pad)])
            for kh in range(ph, 3, 2):
                for kw in range(pw, 3, 2):
                    g.taps.append((0, -((kw - pw) // 2), -((kh - ph) // 2)))
                    g.tapmap.append(kh * 3 + kw)
            out.append((ph, pw, g))
    return out


def geom_patch(N, H, W, C, k) -> ConvGeom:
    """k x k stride-k conv: out (H/k, W/k); tap (kh,kw) has its own strided view."""
    assert H % k == 0 and W % k == 0
    g = ConvGeom(N, H // k, W // k, C)
    for kh in range(k):
        for kw in range(k):
            g.views.append(VqbView(offset=(kh * W + kw) * C, Wv=W // k, Hv=H // k, Nv=N, _pad=0, sw=k * C,
                                   sh=k * W * C, sn=H * W * C))
            g.taps.append((kh * k + kw, 0, 0))
            g.tapmap.append(kh * k + kw)
    return g


# ---- nearest-2x upsample fused into the 3x3 conv (ae.py:164-167), SURVEY.md Appendix A ------------------------------
# out[2a+ph, 2b+pw] = sum_{i,j in 0..1} Wf[ph,pw][i][j] . x[a + OFF[ph][i], b + OFF[pw][j]]
# with Wf[ph,pw][i][j] = sum_{kh in SET[ph][i]} sum_{kw in SET[pw][j]} W[kh,kw]   (4/9 of the MACs, no 4x tensor)
_UP_OFF = {0: (-1, 0), 1: (0, 1)}
_UP_SET = {0: ((0,), (1, 2)), 1: ((0, 1), (2,))}


def _up_mask(ph, pw, i, j) -> int:
    m = 0
    for kh in _UP_SET[ph][i]:
        for kw in _UP_SET[pw][j]:
            m |= 1 << (kh * 3 + kw)
    return m


def geom_up_fwd(N, h, w, C, ph, pw) -> ConvGeom:
    """Phase (ph,pw) of upsample+conv3x3: a 2x2-tap conv over the LOW-RES x writing the (ph,pw) sub-grid of the output."""
    g = ConvGeom(N, h, w, C, [dense_view(N, h, w, C)])
    for i in range(2):
        for j in range(2):
            g.taps.append((0, _UP_OFF[pw][j], _UP_OFF[ph][i]))
            g.tapmask.append(_up_mask(ph, pw, i, j))
    return g


def up_dy_view(N, h, w, Cop, ph, pw) -> VqbView:
    """Parity view (ph,pw) of dy / out [N, 2h, 2w, Cop]."""
    return VqbView(offset=(ph * 2 * w + pw) * Cop, Wv=w, Hv=h, Nv=N, _pad=0, sw=2 * Cop, sh=2 * 2 * w * Cop,
                   sn=4 * h * w * Cop)


def geom_up_dgrad(N, h, w, Cop) -> ConvGeom:
    """dx[a,b] = sum over the 4 phases and 2x2 taps of dy_phase[a - dh, b - dw] . Wf^T : one 16-tap conv over the four
    parity views of dy."""
    g = ConvGeom(N, h, w, Cop)
    for ph in range(2):
        for pw in range(2):
            g.views.append(up_dy_view(N, h, w, Cop, ph, pw))
    for ph in range(2):
        for pw in range(2):
            for i in range(2):
                for j in range(2):
                    g.taps.append((ph * 2 + pw, -_UP_OFF[pw][j], -_UP_OFF[ph][i]))
                    g.tapmask.append(_up_mask(ph, pw, i, j))
    return g


# ---- first-layer "fat pixel" 3x3 conv over an 8-channel image ------------------------------------------------------
# The image lives in a zero-framed buffer [N][H+2][W+2][8]; horizontally adjacent pixels are CONTIGUOUS, so the conv
# becomes 3 taps (kh) whose K run starts at pixel (w-1) and covers the 8 pixels w-1..w+6 = 64 elements = one full
# 128-byte K chunk: columns 0..23 carry the three real taps (kw*8 + c), columns 24..63 meet ZERO weights. The run is 64
# wide (not 24) on purpose: ncu showed the TMA unit spending ~28 cycles per row when the box's inner dimension is
# partly out of range (24 of 64: 640 us for a layer whose HBM floor is 90 us); a fully in-range 128-byte row costs ~1.
# The buffer carries 64 elements of zeroed slack so the last rows stay inside the allocation.
FAT_K = 64


def fat_view(N, H, W) -> VqbView:
    return VqbView(offset=0, Wv=W, Hv=H + 2, Nv=N, _pad=0, sw=8, sh=(W + 2) * 8, sn=(H + 2) * (W + 2) * 8)


def geom_fat3(N, H, W, dgrad=False) -> ConvGeom:
    g = ConvGeom(N, H, W, FAT_K, [fat_view(N, H, W)])
    for kh in range(3):
        g.taps.append((0, 0, kh))
    g.tapmap = [8 - t for t in range(9)] if dgrad else list(range(9))  # 9 packed slots of 8 = 3 fat taps of 24
    return g


def framed_interior_view(N, H, W, C) -> VqbView:
    """The dense (N, H, W) grid seen inside a zero-framed [N][H+2][W+2][C] buffer."""
    return VqbView(offset=((W + 2) + 1) * C, Wv=W, Hv=H, Nv=N, _pad=0, sw=C, sh=(W + 2) * C, sn=(H + 2) * (W + 2) * C)


def conv_desc(g: ConvGeom, Cout: int, out_strides, flags=0, out_f32=False) -> VqbConvDesc:
    """out_strides = (on, oh, ow, oc) in elements."""
    d = VqbConvDesc()
    d.C, d.Cout, d.N, d.H, d.W = g.C, Cout, g.N, g.Ho, g.Wo
    d.nviews, d.ntaps, d.flags, d.out_f32 = len(g.views), len(g.taps), flags, 1 if out_f32 else 0
    d.on, d.oh, d.ow, d.oc = out_strides
    for i, v in enumerate(g.views):
        d.views[i] = v
    for i, (v, dw, dh) in enumerate(g.taps):
        d.taps[i] = VqbTap(view=v, dw=dw, dh=dh, _pad=0)
    return d


def nhwc_strides(H, W, Cs):
    return (H * W * Cs, W * Cs, Cs, 1)


def nchw_strides(C, H, W):
    return (C * H * W, W, 1, H * W)


def wgrad_desc(g: ConvGeom, Cout_pad: int, ksplit: int, dy_view=None, ld_override=0, col_offset=0) -> VqbWgradDesc:
    """x operand geometry = forward geometry g; dy is the dense (N, Ho, Wo, Cout_pad) tensor unless dy_view is given."""
    d = VqbWgradDesc()
    d.C, d.Cout, d.N, d.H, d.W = g.C, Cout_pad, g.N, g.Ho, g.Wo
    d.nviews, d.ntaps, d.ksplit = len(g.views), len(g.taps), ksplit
    d.ld_override, d.col_offset = ld_override, col_offset
    d.dy_view = dy_view if dy_view is not None else dense_view(g.N, g.Ho, g.Wo, Cout_pad)
    for i, v in enumerate(g.views):
        d.views[i] = v
    for i, (v, dw, dh) in enumerate(g.taps):
        d.taps[i] = VqbTap(view=v, dw=dw, dh=dh, _pad=0)
    return d
