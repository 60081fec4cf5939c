"""torch.autograd.Functions over the native sm_100a kernels (libvqb200.so).

Internal activation format: contiguous bf16 tensors of shape [N, H, W, Cp] (NHWC, Cp = channels padded to a multiple
of 8, pad channels zero). Master weights / gradients stay fp32 OIHW nn.Parameters (the reference's state_dict
contract); bf16 packed copies for the tensor-core kernels are caches keyed on the parameter version.

Every op here launches hand-written CUDA; nothing falls back to ATen for the math.
"""
from __future__ import annotations

import math
import os
import threading
import weakref
from typing import Optional

import torch

import native
import plans
from native import EPI_BIAS, EPI_MASK, EPI_RELU, EPI_RES, check, ptr, stream_ptr

_tapmap_cache = {}


def _L():
    return native.load()


def require_cuda(t: torch.Tensor):
    if not t.is_cuda:
        raise RuntimeError("vqgan-training_b200: the hot path runs on sm_100a CUDA tensors only (no CPU fallback)")


def tapmap_tensor(tapmap, device) -> torch.Tensor:
    key = (tuple(tapmap), str(device))
    t = _tapmap_cache.get(key)
    if t is None:
        t = torch.tensor(list(tapmap), dtype=torch.int32, device=device)
        _tapmap_cache[key] = t
    return t


def pack_weights(weight: torch.Tensor, tapmap, transpose: bool, Kpad: int, fold: bool = False) -> torch.Tensor:
    """OIHW fp32 -> bf16 [R][len(tapmap)][Kpad] (R = Cin if transpose else Cout). fold: tapmap holds bit masks of taps
    whose weights are summed (fp32) before the single bf16 rounding."""
    Cout, Cin, KH, KW = weight.shape
    R = Cin if transpose else Cout
    out = torch.empty(R, len(tapmap), Kpad, device=weight.device, dtype=torch.bfloat16)
    tm = tapmap_tensor(tapmap, weight.device)
    w = weight.detach()
    if w.dtype != torch.float32 or not w.is_contiguous():
        w = w.float().contiguous()
    fn = _L().vqb_pack_weights_fold if fold else _L().vqb_pack_weights
    check(fn(ptr(w), ptr(out), Cout, Cin, KH * KW, len(tapmap), ptr(tm), 1 if transpose else 0, Kpad, stream_ptr()),
          "pack_weights")
    return out


class _PackEntry:
    """One cached bf16 GEMM operand of one fp32 OIHW parameter + the recipe to rebuild it in place."""

    __slots__ = ("wref", "out", "tm", "spec", "ver", "__weakref__")

    def __init__(self, weight, out, tm, spec):
        self.wref = weakref.ref(weight)
        self.out, self.tm, self.spec = out, tm, spec
        self.ver = (weight._version, weight.data_ptr())

    def job(self, first_block):
        w = self.wref()
        Cout, Cin, T, nslots, transpose, Kpad, fold, sg, ld_g, ld_r = self.spec
        return native.VqbPackJob(w=w.data_ptr(), out=self.out.data_ptr(), tapmap=self.tm.data_ptr(), Cout=Cout, Cin=Cin,
                                 T=T, nslots=nslots, transpose=transpose, Kpad=Kpad, fold=fold, sg=sg, ld_g=ld_g,
                                 ld_r=ld_r, first_block=first_block, _pad=0)

    def blocks(self):  # one block = an 8-row x 64-k tile, all slots (csrc/optim.cu)
        Cout, Cin, T, nslots, transpose, Kpad = self.spec[:6]
        assert T <= 16
        return -(-(Cin if transpose else Cout) // 8) * -(-Kpad // 64)


# every live pack entry, by the data_ptr of the fp32 master weight it was packed from (weak: caches own the entries)
_pack_registry = {}
_pack_tables = {}


def _new_pack_entry(weight, tapmap, transpose, Kpad, fold, fat=False) -> _PackEntry:
    Cout, Cin, KH, KW = weight.shape
    R = Cin if transpose else Cout
    nslots = len(tapmap)
    if fat:  # [R][9 slots][8] -> [R][3][64]: columns kw*8 + c of each kh row, zero beyond 24 (plans.geom_fat3)
        assert nslots == 9 and Kpad == 8
        out = torch.zeros(R, 3, plans.FAT_K, device=weight.device, dtype=torch.bfloat16)
        spec = (Cout, Cin, KH * KW, nslots, 1 if transpose else 0, Kpad, 0, 3, plans.FAT_K, 3 * plans.FAT_K)
    else:
        out = torch.empty(R, nslots, Kpad, device=weight.device, dtype=torch.bfloat16)
        spec = (Cout, Cin, KH * KW, nslots, 1 if transpose else 0, Kpad, 1 if fold else 0, nslots, 0, nslots * Kpad)
    if weight.dtype != torch.float32 or not weight.is_contiguous():
        raise RuntimeError("vqgan-training_b200: conv master weights must be contiguous fp32 OIHW tensors")
    ent = _PackEntry(weight, out, tapmap_tensor(tapmap, weight.device), spec)
    _pack_registry.setdefault(weight.data_ptr(), weakref.WeakSet()).add(ent)
    return ent


def _run_pack(entries):
    """Re-packs `entries` (in place) with ONE vqb_pack_weights_multi launch; the device job table is cached per entry set."""
    if not entries:
        return
    key = tuple(id(e) for e in entries)
    tab = _pack_tables.get(key)
    if tab is None or any(r() is None for r in tab[3]):
        jobs, nb = [], 0
        for e in entries:
            jobs.append(e.job(nb))
            nb += e.blocks()
        arr = (native.VqbPackJob * len(jobs))(*jobs)
        host = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8)
        dev = host.to(entries[0].out.device)
        if len(_pack_tables) > 64:
            _pack_tables.clear()
        tab = (dev, len(jobs), nb, [weakref.ref(e) for e in entries])
        _pack_tables[key] = tab
    check(_L().vqb_pack_weights_multi(tab[0].data_ptr(), tab[1], tab[2], stream_ptr()), "pack_weights_multi")
    for e in entries:
        w = e.wref()
        e.ver = (w._version, w.data_ptr())


def weights_updated(params=None):
    """Call after master weights changed through a path that does not bump `Tensor._version` (fused / foreach optimizers,
    `.data` writes such as a DDP broadcast): re-packs every cached bf16 operand of `params` (all parameters if None) in
    one launch. A global optimizer post-step hook (registered below) calls this for every torch.optim optimizer."""
    if params is None:
        ptrs = list(_pack_registry.keys())
    else:
        ptrs = [p.data_ptr() for p in params]
    entries = []
    for dp in ptrs:
        ws = _pack_registry.get(dp)
        if not ws:
            continue
        for e in list(ws):
            w = e.wref()
            if w is None or w.data_ptr() != dp:
                ws.discard(e)
                continue
            entries.append(e)
        if not ws:
            _pack_registry.pop(dp, None)
    if entries:
        entries.sort(key=id)
        with torch.no_grad():
            _run_pack(entries)


def _optimizer_post_step(optimizer, args, kwargs):
    try:
        params = [p for g in optimizer.param_groups for p in g["params"] if p.is_cuda]
    except Exception:  # pragma: no cover
        params = None
    if params is None or params:
        weights_updated(params)


try:  # fused AdamW (and any .data-style update) does not bump _version: never trust the version alone
    from torch.optim.optimizer import register_optimizer_step_post_hook as _reg_hook

    _reg_hook(_optimizer_post_step)
except Exception:  # pragma: no cover
    pass


class PackedCache:
    """Per-conv-layer caches: bf16 packed copies of the fp32 OIHW parameter and the shape-dependent geometry objects /
    C descriptors (built once per input shape). A packed copy is valid while (parameter version, data_ptr) are unchanged;
    updates that bypass the version counter are announced through `weights_updated` (global optimizer post-step hook)."""

    def __init__(self):
        self._store = {}
        self._geoms = {}

    def geom(self, key, builder):
        g = self._geoms.get(key)
        if g is None:
            g = builder()
            self._geoms[key] = g
        return g

    def get(self, weight: torch.Tensor, key, tapmap, transpose, Kpad, fold=False, fat=False):
        ent = self._store.get(key)
        if ent is None or ent.wref() is None or ent.ver[1] != weight.data_ptr() or \
                tuple(ent.out.shape[:1]) != ((weight.shape[1] if transpose else weight.shape[0]),):
            ent = _new_pack_entry(weight, tapmap, transpose, Kpad, fold, fat)
            self._store[key] = ent
            with torch.no_grad():
                _run_pack([ent])
        elif ent.ver[0] != weight._version:
            with torch.no_grad():
                _run_pack([ent])
        return ent.out


# Gradient slots: when a parameter lives in a flat.FlatParams store, the kernels that produce its gradient write straight
# into the matching slot of the store's flat gradient buffer and return that view; autograd adopts it as `param.grad`
# (no per-tensor gradient allocation, no copy-in before the all-reduce / fused optimizer).
_grad_slots = {}


def register_grad_slot(param, store, index):
    _grad_slots[param.data_ptr()] = (weakref.ref(store), index)


def grad_out(param: torch.Tensor) -> torch.Tensor:
    """fp32 destination for the gradient of `param`: its flat-store slot (first contribution of this accumulation
    window) or a fresh tensor."""
    ent = _grad_slots.get(param.data_ptr())
    if ent is not None:
        store = ent[0]()
        if store is None:
            _grad_slots.pop(param.data_ptr(), None)
        else:
            q = store.plist[ent[1]]
            if q.data_ptr() == param.data_ptr() and q.shape == param.shape:
                s = store.take_slot(ent[1])
                if s is not None:
                    return s
    return torch.empty(param.shape, device=param.device, dtype=torch.float32)


def _wgrad_block_n(cols: int) -> int:
    for c in (256, 192, 128, 64):
        if cols % c == 0:
            return c
    return 64


def choose_ksplit(g: plans.ConvGeom, Cout_pad: int) -> int:
    """Split-K factor of the weight-gradient GEMM: the split count whose (tile, split) unit count best fills ONE wave of
    the 148 persistent CTAs (measured: one full wave beats two). Mirrors the tile shape rules of csrc/wgrad_gemm.cu
    (256-row tiles + 64-pixel K blocks when Cout >= 256, else 128/128)."""
    cols = len(g.taps) * ((g.C + 63) // 64) * 64
    big = Cout_pad >= 256 and g.C % 64 == 0 and Cout_pad % 64 == 0
    rows = 256 if big else 128
    kpix = 64 if big else 128
    tiles = ((Cout_pad + rows - 1) // rows) * (cols // _wgrad_block_n(cols))

    def p2(v, cap):
        p = 1
        while p < v:
            p <<= 1
        return min(p, cap)

    bw = p2(g.Wo, kpix)
    bh = p2(g.Ho, kpix // bw)
    bn = kpix // (bw * bh)
    boxes = -(-g.Wo // bw) * -(-g.Ho // bh) * -(-g.N // bn)
    # pick the split count whose (tile, split) unit count best fills ONE wave of 148 persistent CTAs: every unit pays a
    # fixed pipeline-fill + fp32-partial-tile epilogue, and the reduction kernel reads ksplit partials, so a single
    # full wave beats two (measured, tools/gpu_probe.py wgbench: 512->512 @ 32^2 ks=4 1471 vs ks=8 1305 TFLOP/s;
    # 128->128 @ 256^2 ks=24 1052 vs ks=49 978; 256->256 @ 128^2 ks=16 1598 vs ks=32 1456)
    sms = 148
    max_ks = max(1, min(boxes // 4, 128, (256 << 20) // max(1, Cout_pad * cols * 4)))
    best, best_score = 1, -1.0
    for ks in range(1, max_ks + 1):
        units = tiles * ks
        waves = -(-units // sms)
        eff = units / (waves * sms)
        score = eff - 0.06 * (waves - 1) - 0.0005 * ks
        if score > best_score:
            best, best_score = ks, score
    return best


class GnLink:
    """Connects a GroupNorm(+swish) forward to the conv that consumes its output, so that the conv's data-gradient
    launch can accumulate the GroupNorm-backward statistics in its epilogue (vqb_conv_gemm_gnbwd) and the GroupNorm
    backward can skip its reduction pass. Filled by GroupNormSiLUFn.forward (x, mr, gamma, beta, groups); `sums` is set by
    ConvFn.backward as (cs, dy data_ptr, dy version, strong ref to dy) and consumed once by GroupNormSiLUFn.backward."""

    __slots__ = ("x", "mr", "gamma", "beta", "groups", "silu", "sums")

    def __init__(self):
        self.x = self.mr = self.gamma = self.beta = None
        self.groups, self.silu, self.sums = 0, False, None


# Opt-in (VQB_GN_BWD_FUSE=1): measured on the B=32 step, the extra ~1200 instructions per 64-channel group in the four
# epilogue warps are NOT hidden behind the main loop of the 128-channel layers (conv_gemm total 46.1 -> 64.9 ms for a
# 5.2 ms saving in gn_bwd_reduce); kept, with its parity test, for layers with long K loops.
_GN_BWD_FUSE = os.environ.get("VQB_GN_BWD_FUSE", "0") == "1"
_GN_BWD_FUSE_MIN_C = int(os.environ.get("VQB_GN_BWD_FUSE_MIN_C", "0"))


def conv_gnbwd_supported(g: plans.ConvGeom, Cout: int, out_strides, groups: int) -> bool:
    descs = g.__dict__.setdefault("_descs", {})
    key = ("gnbwd_ok", Cout, tuple(out_strides), groups)
    ok = descs.get(key)
    if ok is None:
        ok = bool(_L().vqb_conv_gnbwd_ok(plans.conv_desc(g, Cout, out_strides, 0, False), groups))
        descs[key] = ok
    return ok


gnbwd_fused_launches = 0  # how many data-gradient launches carried fused GroupNorm-backward statistics (tests)


def run_conv_gemm_gnbwd(g: plans.ConvGeom, a, wp, Cout, out, out_strides, link: "GnLink", cs):
    global gnbwd_fused_launches
    gnbwd_fused_launches += 1
    dk = (Cout, tuple(out_strides), 0, False)
    descs = g.__dict__.setdefault("_descs", {})
    d = descs.get(dk)
    if d is None:
        d = plans.conv_desc(g, Cout, out_strides, 0, False)
        descs[dk] = d
    ga, be = link.gamma.detach(), link.beta.detach()
    fuse = native.VqbGnBwdFuse(x=ptr(link.x), mr=ptr(link.mr), gamma=ptr(ga), beta=ptr(be), cs=ptr(cs),
                               groups=link.groups, _pad=0)
    check(_L().vqb_conv_gemm_gnbwd(d, ptr(a), ptr(wp), 0, ptr(out), fuse, stream_ptr()), "conv_gemm_gnbwd")


def run_conv_gemm(g: plans.ConvGeom, a: torch.Tensor, wp: torch.Tensor, Cout: int, out: torch.Tensor, out_strides,
                  out_ptr_offset_bytes=0, bias=None, res=None, mask=None, relu=False, out_f32=False, stats=None):
    flags = (EPI_BIAS if bias is not None else 0) | (EPI_RES if res is not None else 0) | \
            (EPI_MASK if mask is not None else 0) | (EPI_RELU if relu else 0) | \
            (native.EPI_STATS if stats is not None else 0)
    dk = (Cout, tuple(out_strides), flags, out_f32)
    descs = g.__dict__.setdefault("_descs", {})
    d = descs.get(dk)
    if d is None:
        d = plans.conv_desc(g, Cout, out_strides, flags, out_f32)
        descs[dk] = d
    off = out_ptr_offset_bytes
    check(_L().vqb_conv_gemm(d, ptr(a), ptr(wp), ptr(bias), (ptr(res) + off) if res is not None else 0,
                             (ptr(mask) + off) if mask is not None else 0, ptr(out) + off, ptr(stats), stream_ptr()),
          "conv_gemm")


def conv_stats_supported(g: plans.ConvGeom, Cout: int, out_strides) -> bool:
    """Can the conv epilogue produce the GroupNorm statistics of its output for this geometry? (cached per geometry)"""
    descs = g.__dict__.setdefault("_descs", {})
    key = ("stats_ok", Cout, tuple(out_strides))
    ok = descs.get(key)
    if ok is None:
        ok = bool(_L().vqb_conv_stats_ok(plans.conv_desc(g, Cout, out_strides, 0, False)))
        descs[key] = ok
    return ok


def run_wgrad(g: plans.ConvGeom, x: torch.Tensor, dy: torch.Tensor, weight_shape, Cout_pad: int,
              dy_view=None, out=None) -> torch.Tensor:
    """-> OIHW fp32 gradient for a conv whose forward geometry is g (weight_shape = (Cout, K per tap, taps_h, taps_w))."""
    Cout, Cin, KH, KW = weight_shape
    wk = ("wgrad", Cout_pad, dy_view is not None)
    descs = g.__dict__.setdefault("_descs", {})
    ent = descs.get(wk)
    if ent is None:
        ksplit = choose_ksplit(g, Cout_pad)
        ent = (ksplit, plans.wgrad_desc(g, Cout_pad, ksplit, dy_view=dy_view), _L().vqb_wgrad_cols(len(g.taps), g.C))
        descs[wk] = ent
    ksplit, d, cols = ent
    partial = torch.empty(ksplit, Cout_pad, cols, device=x.device, dtype=torch.float32)
    check(_L().vqb_wgrad_gemm(d, ptr(dy), ptr(x), ptr(partial), stream_ptr()), "wgrad_gemm")
    grad = out if out is not None else torch.empty(Cout, Cin, KH, KW, device=x.device, dtype=torch.float32)
    assert grad.shape == (Cout, Cin, KH, KW) and grad.is_contiguous()
    tm = tapmap_tensor(g.tapmap if len(g.tapmap) == len(g.taps) else list(range(len(g.taps))), x.device)
    check(_L().vqb_wgrad_reduce(ptr(partial), ptr(grad), ksplit, Cout, Cout_pad, Cin, KH * KW, len(g.taps),
                                cols // len(g.taps), ptr(tm), 0, stream_ptr()), "wgrad_reduce")
    return grad


# One-slot side channel from GroupNormSiLUFn.backward to the backward of the conv that produced the normalised tensor:
# the GN backward apply pass already streams dx (= that conv's dy), so it also emits the per-channel sums (= the conv's
# bias gradient). The slot holds a strong reference to dx, so a matching data_ptr can only be that very tensor; the
# version check rejects a tensor that autograd accumulated into in place. Any mismatch falls back to vqb_colsum.
class _Slot(threading.local):
    """one slot per thread: autograd runs a device's backward on one worker thread, so a re-entrant or concurrent
    backward on another thread can neither see nor clobber this one's hand-over"""

    def __init__(self):
        self.v = None

    def __getitem__(self, i):
        return self.v

    def __setitem__(self, i, val):
        self.v = val


_dx_colsum_slot = _Slot()


def _take_dx_colsum(dy: torch.Tensor, C: int):
    ent = _dx_colsum_slot[0]
    if ent is None:
        return None
    t, ver, cs = ent
    if (t.data_ptr() == dy.data_ptr() and t.shape == dy.shape and t.stride() == dy.stride() and t.dtype == dy.dtype
            and dy._version == ver and cs.numel() == C):
        _dx_colsum_slot[0] = None
        return cs
    return None


def colsum(x2d_rows: int, x: torch.Tensor, C: int, out=None) -> torch.Tensor:
    if out is None:
        out = torch.empty(C, device=x.device, dtype=torch.float32)
    check(_L().vqb_colsum(ptr(x), ptr(out), x2d_rows, C, stream_ptr()), "colsum")
    return out


# ----------------------------------------------------------------------------------------------------------------------
class ToNHWC(torch.autograd.Function):
    """[N,C,H,W] fp32 -> [N,H,W,Cp] bf16, optional per-channel (x - shift) * inv_scale (LPIPS ScalingLayer,
    utils.py:70-71). Backward: NHWC bf16 grad -> NCHW fp32 (* inv_scale)."""

    @staticmethod
    def forward(ctx, x, shift, inv_scale, frame):
        require_cuda(x)
        x = x.detach()
        if x.dtype != torch.float32 or not x.is_contiguous():
            x = x.float().contiguous()
        N, C, H, W = x.shape
        Cp = plans.cpad(C)
        if frame:  # zero-framed [N, H+2, W+2, Cp] for the "fat pixel" first-layer conv
            y = alloc_framed(N, H, W, Cp, x.device)
            check(_L().vqb_nchw_to_nhwc_pad(ptr(x), ptr(y), N, C, H, W, Cp, 1, ptr(shift), ptr(inv_scale),
                                            stream_ptr()), "nchw_to_nhwc_pad")
        else:
            y = torch.empty(N, H, W, Cp, device=x.device, dtype=torch.bfloat16)
            check(_L().vqb_nchw_to_nhwc(ptr(x), ptr(y), N, C, H, W, Cp, ptr(shift), ptr(inv_scale), stream_ptr()),
                  "nchw_to_nhwc")
        ctx.shape = (N, C, H, W, Cp)
        ctx.inv_scale, ctx.frame = inv_scale, frame
        return y

    @staticmethod
    def backward(ctx, gy):
        N, C, H, W, Cp = ctx.shape
        gy = gy.contiguous()
        gx = torch.empty(N, C, H, W, device=gy.device, dtype=torch.float32)
        if ctx.frame:
            check(_L().vqb_nhwc_to_nchw_pad(ptr(gy), ptr(gx), N, C, H, W, Cp, 1, ptr(ctx.inv_scale), stream_ptr()),
                  "nhwc_to_nchw_pad")
        else:
            check(_L().vqb_nhwc_to_nchw(ptr(gy), ptr(gx), N, C, H, W, Cp, ptr(ctx.inv_scale), stream_ptr()),
                  "nhwc_to_nchw")
        return gx, None, None, None


def alloc_framed(N, H, W, C, device) -> torch.Tensor:
    """Zeroed [N, H+2, W+2, C] bf16 buffer followed by 64 elements of zeroed slack (fat-pixel K runs read up to 5 pixels
    past the last one; they meet zero weights but must stay inside the allocation and finite)."""
    n = N * (H + 2) * (W + 2) * C
    return torch.zeros(n + 64, device=device, dtype=torch.bfloat16)[:n].view(N, H + 2, W + 2, C)


def _fat_weights(cache: "PackedCache", weight, key, tapmap, transpose, Kpad):
    """[R][9 slots][8] packing laid out as [R][3][64]: columns kw*8 + c of each kh row, zero beyond 24 (plans.geom_fat3)."""
    return cache.get(weight, tuple(key) + ("k64",), tapmap, transpose, Kpad, fat=True)


_fat_state = {"ok": None}


def fat_conv_enabled() -> bool:
    """One-time self check of the fat-pixel first-layer path (it relies on a TMA map whose pixel stride (16 B) is smaller
    than its 48-byte inner extent): run a tiny conv both ways; disable the path on any error or mismatch."""
    import os

    if os.environ.get("VQB_FAT_CONV", "1") != "1":
        return False
    if _fat_state["ok"] is None:
        _fat_state["ok"] = False
        try:
            g = torch.Generator(device="cuda").manual_seed(1)
            x = torch.rand(2, 3, 16, 24, device="cuda", generator=g) - 0.5
            w = torch.rand(64, 3, 3, 3, device="cuda", generator=g) - 0.5
            c1, c2 = PackedCache(), PackedCache()
            a = conv(ToNHWC.apply(x, None, None, False), w, None, c1, "s1")
            b = conv(ToNHWC.apply(x, None, None, True), w, None, c2, "fat3")
            torch.cuda.synchronize()
            _fat_state["ok"] = bool(torch.allclose(a.float(), b.float(), rtol=2e-2, atol=2e-2))
        except Exception:
            _fat_state["ok"] = False
    return _fat_state["ok"]


def wavelet_to_nhwc(x: torch.Tensor, filt: torch.Tensor) -> torch.Tensor:
    """[N,C,H,W] fp32 image -> [N,H/2,W/2,cpad(4C)] bf16: the wavelet front-end (utils.py:229-247) fused with the layout
    conversion. Input-side op: the image is data, so there is no backward."""
    require_cuda(x)
    if x.requires_grad:
        raise RuntimeError("wavelet front-end: the input image must not require grad (input-side op without backward)")
    x = x.detach().float().contiguous()
    N, C, H, W = x.shape
    Cp = plans.cpad(4 * C)
    y = torch.empty(N, H // 2, W // 2, Cp, device=x.device, dtype=torch.bfloat16)
    f = filt.detach().to(device=x.device, dtype=torch.float32).reshape(4, 36).contiguous()
    check(_L().vqb_wavelet_fwd(ptr(x), ptr(y), ptr(f), N, C, H, W, Cp, stream_ptr()), "wavelet_fwd")
    return y


def to_nhwc(x, shift=None, inv_scale=None, frame=False):
    return ToNHWC.apply(x, shift, inv_scale, frame)


class ToNCHW(torch.autograd.Function):
    """[N,H,W,Cp] bf16 -> [N,C,H,W] fp32 (module-boundary output when a caller wants the reference layout)."""

    @staticmethod
    def forward(ctx, y, C):
        N, H, W, Cp = y.shape
        y = y.contiguous()
        x = torch.empty(N, C, H, W, device=y.device, dtype=torch.float32)
        check(_L().vqb_nhwc_to_nchw(ptr(y), ptr(x), N, C, H, W, Cp, 0, stream_ptr()), "nhwc_to_nchw")
        ctx.shape = (N, C, H, W, Cp)
        return x

    @staticmethod
    def backward(ctx, gx):
        N, C, H, W, Cp = ctx.shape
        gx = gx.float().contiguous()
        gy = torch.empty(N, H, W, Cp, device=gx.device, dtype=torch.bfloat16)
        check(_L().vqb_nchw_to_nhwc(ptr(gx), ptr(gy), N, C, H, W, Cp, 0, 0, stream_ptr()), "nchw_to_nhwc")
        return gy, None


def to_nchw(y, C):
    return ToNCHW.apply(y, C)


# ----------------------------------------------------------------------------------------------------------------------
class ConvFn(torch.autograd.Function):
    """Convolution through the tcgen05 implicit-GEMM kernel.

    kind: "s1" (k x k stride 1 same), "s2" (Downsample: pad (0,1,0,1) + 3x3 stride 2), "patch" (k x k stride k).
    opts: relu (fused ReLU epilogue; the incoming gradient is then expected to be already gated by out > 0, which every
    consumer of a ReLU output in this package does), input_is_relu (gate the data gradient by x > 0 in the dgrad
    epilogue), nchw_out (write fp32 [N,Cout,H,W] directly: encoder z / decoder image)."""

    @staticmethod
    def forward(ctx, x, weight, bias, residual, cache, kind, relu, input_is_relu, nchw_out, want_stats=False,
                gn_link=None):
        require_cuda(x)
        N, H, W, Cp = x.shape
        Cout, Cin, KH, KW = weight.shape
        assert Cp == plans.cpad(Cin), f"conv input has {Cp} channels, weight expects {Cin}"
        x = x.contiguous()
        if kind == "fat3":  # x is the zero-framed [N, H+2, W+2, 8] image
            assert Cp == 8 and KH == 3
            H, W = H - 2, W - 2
            g = cache.geom(("fat", N, H, W), lambda: plans.geom_fat3(N, H, W))
        elif kind == "s1":
            g = cache.geom(("f", N, H, W), lambda: plans.geom_s1(N, H, W, Cp, KH))
        elif kind == "s2":
            g = cache.geom(("f", N, H, W), lambda: plans.geom_s2(N, H, W, Cp))
        elif kind == "patch":
            g = cache.geom(("f", N, H, W), lambda: plans.geom_patch(N, H, W, Cp, KH))
        else:
            raise ValueError(kind)
        if kind == "fat3":
            wp = _fat_weights(cache, weight, ("fwd", kind), g.tapmap, False, Cp)
        else:
            wp = cache.get(weight, ("fwd", kind), g.tapmap, False, Cp)
        Cop = plans.cpad(Cout)
        ctx.HW = (H, W)
        b = None
        if bias is not None:
            b = bias.detach()
            if b.dtype != torch.float32:
                b = b.float()
        if nchw_out:
            out = torch.empty(N, Cout, g.Ho, g.Wo, device=x.device, dtype=torch.float32)
            run_conv_gemm(g, x, wp, Cout, out, plans.nchw_strides(Cout, g.Ho, g.Wo), bias=b, relu=relu, out_f32=True)
        else:
            alloc = torch.empty if Cop == Cout else torch.zeros
            out = alloc(N, g.Ho, g.Wo, Cop, device=x.device, dtype=torch.bfloat16)
            res = residual.contiguous() if residual is not None else None
            ostr = plans.nhwc_strides(g.Ho, g.Wo, Cop)
            stats = None
            if want_stats and Cop == Cout and conv_stats_supported(g, Cout, ostr):
                stats = torch.zeros(N, Cout, 2, device=x.device, dtype=torch.float32)
            run_conv_gemm(g, x, wp, Cout, out, ostr, bias=b, res=res, relu=relu, stats=stats)
        ctx.save_for_backward(x, weight)
        ctx.cache, ctx.kind, ctx.g = cache, kind, g
        ctx.has_bias, ctx.has_res = bias is not None, residual is not None
        ctx.input_is_relu, ctx.nchw_out = input_is_relu, nchw_out
        ctx.gn_link = gn_link if _GN_BWD_FUSE else None
        if want_stats and not nchw_out:
            if stats is None:
                stats = torch.empty(0, device=x.device)  # "not available" marker
            ctx.mark_non_differentiable(stats)
            return out, stats
        return out

    @staticmethod
    def backward(ctx, gout, _gstats=None):
        x, weight = ctx.saved_tensors
        g, kind, cache = ctx.g, ctx.kind, ctx.cache
        N, _, _, Cp = x.shape
        H, W = ctx.HW
        Cout, Cin, KH, KW = weight.shape
        Cop = plans.cpad(Cout)
        dy_framed = False
        if ctx.nchw_out:
            gn = gout.float().contiguous()
            if Cop == 8 and KH == 3 and kind == "s1" and fat_conv_enabled():
                # tiny-Cout conv (decoder conv_out): keep dy in a zero-framed buffer so that the data gradient runs as a
                # 3-tap fat-pixel conv (24-wide K runs) instead of 9 taps of 8 channels
                dy_framed = True
                dy = alloc_framed(N, g.Ho, g.Wo, Cop, x.device)
                check(_L().vqb_nchw_to_nhwc_pad(ptr(gn), ptr(dy), N, Cout, g.Ho, g.Wo, Cop, 1, 0, 0, stream_ptr()),
                      "nchw_to_nhwc_pad")
            else:
                dy = torch.empty(N, g.Ho, g.Wo, Cop, device=x.device, dtype=torch.bfloat16)
                check(_L().vqb_nchw_to_nhwc(ptr(gn), ptr(dy), N, Cout, g.Ho, g.Wo, Cop, 0, 0, stream_ptr()),
                      "nchw_to_nhwc")
        else:
            dy = gout.contiguous()
        gx = gw = gb = gres = None
        if ctx.needs_input_grad[0]:
            mask = x if ctx.input_is_relu else None
            gx_alloc = torch.empty if Cp == Cin else torch.zeros
            if kind == "fat3":  # gradient w.r.t. the framed image: write the interior of a zero-framed buffer
                gx = torch.zeros(N, H + 2, W + 2, Cp, device=x.device, dtype=torch.bfloat16)
                gd = cache.geom(("d", N, H, W), lambda: plans.geom_s1_dgrad(N, H, W, Cop, KH))
                wpd = cache.get(weight, ("dgrad", "s1"), gd.tapmap, True, Cop)
                run_conv_gemm(gd, dy, wpd, Cin, gx, ((H + 2) * (W + 2) * Cp, (W + 2) * Cp, Cp, 1),
                              out_ptr_offset_bytes=((W + 2) + 1) * Cp * 2)
            elif dy_framed:
                gx = gx_alloc(N, H, W, Cp, device=x.device, dtype=torch.bfloat16)
                gdf = cache.geom(("dfat", N, H, W), lambda: plans.geom_fat3(N, H, W, dgrad=True))
                wpd = _fat_weights(cache, weight, ("dgrad", "fat3"), gdf.tapmap, True, Cop)
                run_conv_gemm(gdf, dy, wpd, Cin, gx, plans.nhwc_strides(H, W, Cp), mask=mask)
            else:
                gx = gx_alloc(N, H, W, Cp, device=x.device, dtype=torch.bfloat16)
            if kind == "fat3" or dy_framed:
                pass
            elif kind == "s1":
                gd = cache.geom(("d", N, H, W), lambda: plans.geom_s1_dgrad(N, H, W, Cop, KH))
                wpd = cache.get(weight, ("dgrad", kind), gd.tapmap, True, Cop)
                link = ctx.gn_link
                ostr = plans.nhwc_strides(H, W, Cp)
                if (link is not None and mask is None and Cp == Cin and Cin >= _GN_BWD_FUSE_MIN_C and link.silu
                        and link.x is not None
                        and link.x.shape == gx.shape and conv_gnbwd_supported(gd, Cin, ostr, link.groups)):
                    # this data gradient IS the dy of the GroupNorm(+swish) that produced x: its epilogue also
                    # accumulates that GroupNorm's backward statistics (the separate reduction pass disappears)
                    cs = torch.zeros(N, Cin, 2, device=x.device, dtype=torch.float32)
                    run_conv_gemm_gnbwd(gd, dy, wpd, Cin, gx, ostr, link, cs)
                    link.sums = (cs, gx.data_ptr(), gx._version, gx)
                else:
                    run_conv_gemm(gd, dy, wpd, Cin, gx, ostr, mask=mask)
            elif kind == "s2":
                for ph, pw, gd in cache.geom(("d", N, H, W), lambda: plans.geom_s2_dgrad_classes(N, H, W, Cop)):
                    wpd = cache.get(weight, ("dgrad", kind, ph, pw), gd.tapmap, True, Cop)
                    run_conv_gemm(gd, dy, wpd, Cin, gx, (H * W * Cp, 2 * W * Cp, 2 * Cp, 1),
                                  out_ptr_offset_bytes=(ph * W + pw) * Cp * 2, mask=mask)
            elif kind == "patch":
                # non-overlapping windows: each input pixel belongs to exactly one (output pixel, tap): one 1-tap
                # "conv" per tap writing the strided sub-grid of dx
                k = KH
                for kh in range(k):
                    for kw in range(k):
                        gd = cache.geom(("d", N, H, W, kh, kw), lambda: plans.ConvGeom(
                            N, g.Ho, g.Wo, Cop, [native.dense_view(N, g.Ho, g.Wo, Cop)], [(0, 0, 0)], [kh * k + kw]))
                        wpd = cache.get(weight, ("dgrad", kind, kh, kw), gd.tapmap, True, Cop)
                        run_conv_gemm(gd, dy, wpd, Cin, gx, (H * W * Cp, k * W * Cp, k * Cp, 1),
                                      out_ptr_offset_bytes=(kh * W + kw) * Cp * 2, mask=mask)
        if ctx.needs_input_grad[1]:
            if kind == "fat3":  # [Cout][kw*8 + c][kh] -> OIHW
                g3 = run_wgrad(g, x, dy, (Cout, plans.FAT_K, 3, 1), Cop)  # [Cout][kw*8 + c (24 real of 64)][kh]
                gw = g3[:, :24, :, 0].reshape(Cout, 3, 8, 3)[:, :, :Cin, :].permute(0, 2, 3, 1).contiguous()
            elif dy_framed:
                gw = run_wgrad(g, x, dy, weight.shape, Cop,
                               dy_view=plans.framed_interior_view(N, g.Ho, g.Wo, Cop), out=grad_out(weight))
            else:
                gw = run_wgrad(g, x, dy, weight.shape, Cop, out=grad_out(weight))
        if ctx.has_bias and ctx.needs_input_grad[2]:
            rows = N * (g.Ho + 2) * (g.Wo + 2) if dy_framed else N * g.Ho * g.Wo  # the zero frame adds nothing
            gb = None if dy_framed else _take_dx_colsum(dy, Cop)
            if gb is None:
                gb = colsum(rows, dy, Cop)[:Cout]
        if ctx.has_res and ctx.needs_input_grad[3]:
            gres = dy
        return gx, gw, gb, gres, None, None, None, None, None, None, None


def conv(x, weight, bias, cache, kind="s1", residual=None, relu=False, input_is_relu=False, nchw_out=False,
         want_stats=False, gn_link=None):
    """-> out, or (out, stats) when want_stats (stats is None if the epilogue cannot produce them for this shape).
    gn_link: the GnLink of the GroupNorm(+swish) whose output `x` is (fused GroupNorm-backward statistics)."""
    if want_stats and not nchw_out:
        out, st = ConvFn.apply(x, weight, bias, residual, cache, kind, relu, input_is_relu, nchw_out, True, gn_link)
        return out, (st if st.numel() > 0 else None)
    return ConvFn.apply(x, weight, bias, residual, cache, kind, relu, input_is_relu, nchw_out, False, gn_link)


# ----------------------------------------------------------------------------------------------------------------------
class UpConvFn(torch.autograd.Function):
    """Upsample (nearest x2, ae.py:165) + conv3x3 p1 (ae.py:166) without materialising the 4x tensor: four phase convs with
    2x2 folded taps over the low-res input (4/9 of the MACs). Backward: one 16-tap conv over the
    four parity views of dy (data gradient) and four phase weight-gradient GEMMs unfolded by vqb_wgrad_reduce_fold."""

    @staticmethod
    def forward(ctx, x, weight, bias, cache, want_stats=False):
        require_cuda(x)
        x = x.contiguous()
        N, h, w, Cp = x.shape
        Cout, Cin, KH, KW = weight.shape
        assert KH == 3 and KW == 3 and Cp == plans.cpad(Cin)
        Cop = plans.cpad(Cout)
        alloc = torch.empty if Cop == Cout else torch.zeros
        out = alloc(N, 2 * h, 2 * w, Cop, device=x.device, dtype=torch.bfloat16)
        b = bias.detach().float() if bias is not None else None
        strides = (4 * h * w * Cop, 2 * 2 * w * Cop, 2 * Cop, 1)
        stats = None
        g00 = cache.geom(("uf", N, h, w, 0, 0), lambda: plans.geom_up_fwd(N, h, w, Cp, 0, 0))
        if want_stats and Cop == Cout and conv_stats_supported(g00, Cout, strides):
            stats = torch.zeros(N, Cout, 2, device=x.device, dtype=torch.float32)  # the 4 phase launches accumulate
        for ph in range(2):
            for pw in range(2):
                g = cache.geom(("uf", N, h, w, ph, pw), lambda: plans.geom_up_fwd(N, h, w, Cp, ph, pw))
                wp = cache.get(weight, ("ufwd", ph, pw), g.tapmask, False, Cp, fold=True)
                run_conv_gemm(g, x, wp, Cout, out, strides, out_ptr_offset_bytes=(ph * 2 * w + pw) * Cop * 2, bias=b,
                              stats=stats)
        ctx.save_for_backward(x, weight)
        ctx.cache, ctx.has_bias = cache, bias is not None
        if want_stats:
            if stats is None:
                stats = torch.empty(0, device=x.device)
            ctx.mark_non_differentiable(stats)
            return out, stats
        return out

    @staticmethod
    def backward(ctx, gout, _gstats=None):
        x, weight = ctx.saved_tensors
        cache = ctx.cache
        N, h, w, Cp = x.shape
        Cout, Cin, KH, KW = weight.shape
        Cop = plans.cpad(Cout)
        dy = gout.contiguous()
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            gd = cache.geom(("ud", N, h, w), lambda: plans.geom_up_dgrad(N, h, w, Cop))
            wpd = cache.get(weight, ("udgrad",), gd.tapmask, True, Cop, fold=True)
            gx_alloc = torch.empty if Cp == Cin else torch.zeros
            gx = gx_alloc(N, h, w, Cp, device=x.device, dtype=torch.bfloat16)
            run_conv_gemm(gd, dy, wpd, Cin, gx, plans.nhwc_strides(h, w, Cp))
        if ctx.needs_input_grad[1]:
            C64 = ((Cp + 63) // 64) * 64
            g00 = cache.geom(("uf", N, h, w, 0, 0), lambda: plans.geom_up_fwd(N, h, w, Cp, 0, 0))
            ksplit = choose_ksplit(g00, Cop)
            partial = torch.empty(ksplit, Cop, 16 * C64, device=x.device, dtype=torch.float32)
            masks = []
            for ph in range(2):
                for pw in range(2):
                    g = cache.geom(("uf", N, h, w, ph, pw), lambda: plans.geom_up_fwd(N, h, w, Cp, ph, pw))
                    dk = ("uwgrad", Cop, ksplit)
                    descs = g.__dict__.setdefault("_descs", {})
                    d = descs.get(dk)
                    if d is None:
                        d = plans.wgrad_desc(g, Cop, ksplit, dy_view=plans.up_dy_view(N, h, w, Cop, ph, pw),
                                             ld_override=16 * C64, col_offset=(ph * 2 + pw) * 4 * C64)
                        descs[dk] = d
                    check(_L().vqb_wgrad_gemm(d, ptr(dy), ptr(x), ptr(partial), stream_ptr()), "wgrad_gemm(up)")
                    masks += g.tapmask
            gw = grad_out(weight)
            tm = tapmap_tensor(masks, x.device)
            check(_L().vqb_wgrad_reduce_fold(ptr(partial), ptr(gw), ksplit, Cout, Cop, Cin, 9, 16, C64, ptr(tm),
                                             stream_ptr()), "wgrad_reduce_fold")
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = _take_dx_colsum(dy, Cop)
            if gb is None:
                gb = colsum(N * 4 * h * w, dy, Cop)[:Cout]
        return gx, gw, gb, None, None


def upsample_conv(x, weight, bias, cache, want_stats=False):
    if want_stats:
        out, st = UpConvFn.apply(x, weight, bias, cache, True)
        return out, (st if st.numel() > 0 else None)
    return UpConvFn.apply(x, weight, bias, cache, False)


_GN_COLSUM = os.environ.get("VQB_GN_COLSUM", "1") == "1"


class GroupNormSiLUFn(torch.autograd.Function):
    """FP32GroupNorm (32 groups, eps 1e-6, biased variance, fp32 statistics; ae.py:41-53) fused with swish
    (ae.py:13-14): one statistics pass + one apply pass over bf16 NHWC, instead of cast/GN/cast/sigmoid/mul.

    with_skip=True additionally returns the input itself as a second output (the ResnetBlock skip connection): the
    gradient arriving through that output is summed into dx INSIDE the backward apply kernel instead of by a separate
    autograd accumulation kernel."""

    @staticmethod
    def forward(ctx, x, gamma, beta, groups, eps, silu, with_skip, chsums=None, link=None):
        require_cuda(x)
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty_like(x)
        mr = torch.empty(N, groups, 2, device=x.device, dtype=torch.float32)
        ga, be = gamma.detach().float(), beta.detach().float()
        if chsums is not None:  # statistics were accumulated by the epilogue of the conv that produced x
            check(_L().vqb_gn_silu_fwd_pre(ptr(x), ptr(y), ptr(ga), ptr(be), ptr(mr), ptr(chsums), N, H * W, C, groups,
                                           eps, 1 if silu else 0, stream_ptr()), "gn_silu_fwd_pre")
        else:
            ws = torch.empty(N * C * 2, device=x.device, dtype=torch.float64)
            check(_L().vqb_gn_silu_fwd(ptr(x), ptr(y), ptr(ga), ptr(be), ptr(mr), ptr(ws), N, H * W, C, groups, eps,
                                       1 if silu else 0, stream_ptr()), "gn_silu_fwd")
        ctx.save_for_backward(x, gamma, beta, mr)
        ctx.groups, ctx.silu, ctx.with_skip = groups, silu, with_skip
        ctx.link = link
        if link is not None:
            link.x, link.mr, link.gamma, link.beta, link.groups, link.silu = x, mr, gamma, beta, groups, bool(silu)
        ctx.set_materialize_grads(False)  # an unused output arrives as None, not as a zero tensor
        if with_skip:
            return y, x.view_as(x)
        return y

    @staticmethod
    def backward(ctx, gy, gskip=None):
        x, gamma, beta, mr = ctx.saved_tensors
        N, H, W, C = x.shape
        if gy is None:  # only the skip output was used
            return (gskip, None, None, None, None, None, None, None, None)
        gy = gy.contiguous()
        add = gskip.contiguous() if gskip is not None else None
        dx = torch.empty_like(x)
        dg, db = grad_out(gamma), grad_out(beta)
        ws = torch.empty(N * C * 2 + N * ctx.groups * 2, device=x.device, dtype=torch.float32)
        ga, be = gamma.detach().float(), beta.detach().float()
        cs = torch.empty(C, device=x.device, dtype=torch.float32) if _GN_COLSUM else None
        pre = None
        link = ctx.link
        if link is not None and link.sums is not None:
            pcs, dptr, dver, dref = link.sums
            link.sums = None
            if dptr == gy.data_ptr() and dref.shape == gy.shape and gy._version == dver and pcs.shape == (N, C, 2):
                pre = pcs  # the consumer conv's data-gradient epilogue already accumulated (sum du, sum du*xhat)
        if link is not None:
            link.x = link.mr = link.gamma = link.beta = None  # drop the references once the backward ran
        if pre is not None:
            check(_L().vqb_gn_silu_bwd_pre(ptr(x), ptr(gy), ptr(add), ptr(dx), ptr(ga), ptr(be), ptr(mr), ptr(pre),
                                           ptr(dg), ptr(db), ptr(ws), N, H * W, C, ctx.groups, 1 if ctx.silu else 0,
                                           ptr(cs), stream_ptr()), "gn_silu_bwd_pre")
        else:
            check(_L().vqb_gn_silu_bwd(ptr(x), ptr(gy), ptr(add), ptr(dx), ptr(ga), ptr(be), ptr(mr), ptr(dg), ptr(db),
                                       ptr(ws), N, H * W, C, ctx.groups, 1 if ctx.silu else 0, ptr(cs), stream_ptr()),
                  "gn_silu_bwd")
        _dx_colsum_slot[0] = (dx, dx._version, cs) if cs is not None else None
        return dx, dg, db, None, None, None, None, None, None


def group_norm_silu(x, gamma, beta, groups=32, eps=1e-6, silu=True, with_skip=False, chsums=None, link=None):
    return GroupNormSiLUFn.apply(x, gamma, beta, groups, eps, silu, with_skip, chsums, link)


class Upsample2xFn(torch.autograd.Function):
    """nearest x2 (ae.py:165); backward = 2x2 sum."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, 2 * H, 2 * W, C, device=x.device, dtype=x.dtype)
        check(_L().vqb_upsample2x_fwd(ptr(x), ptr(y), N, H, W, C, stream_ptr()), "upsample2x_fwd")
        return y

    @staticmethod
    def backward(ctx, gy):
        gy = gy.contiguous()
        N, H2, W2, C = gy.shape
        gx = torch.empty(N, H2 // 2, W2 // 2, C, device=gy.device, dtype=gy.dtype)
        check(_L().vqb_upsample2x_bwd(ptr(gy), ptr(gx), N, H2 // 2, W2 // 2, C, stream_ptr()), "upsample2x_bwd")
        return gx


def upsample2x(x):
    return Upsample2xFn.apply(x)


class MaxPool2Fn(torch.autograd.Function):
    """2x2/2 max-pool of a post-ReLU activation. The backward routes dy to the first maximum of each window and gates
    it by x > 0, i.e. it returns the gradient of the *pre-activation* of the conv that produced x."""

    @staticmethod
    def forward(ctx, x):
        x = x.contiguous()
        N, H, W, C = x.shape
        y = torch.empty(N, H // 2, W // 2, C, device=x.device, dtype=x.dtype)
        check(_L().vqb_maxpool2_fwd(ptr(x), ptr(y), N, H // 2, W // 2, C, stream_ptr()), "maxpool2_fwd")
        ctx.save_for_backward(x)
        return y

    @staticmethod
    def backward(ctx, gy):
        (x,) = ctx.saved_tensors
        gy = gy.contiguous()
        N, H, W, C = x.shape
        gx = torch.empty_like(x)
        check(_L().vqb_maxpool2_bwd(ptr(x), ptr(gy), 0, ptr(gx), N, H // 2, W // 2, C, 1, stream_ptr()),
              "maxpool2_bwd")
        return gx


def maxpool2(x):
    return MaxPool2Fn.apply(x)


class LpipsTailFn(torch.autograd.Function):
    """One LPIPS layer (utils.py:46-53,134-140): unit-normalise over channels, squared difference, [train mode: Dropout(0.5)
    with the counter-based mask of `seed`], 1x1 lin, spatial mean -> [N]. Gradient only w.r.t. f0 (reconstruction branch),
    gated by f0 > 0 (post-ReLU feature)."""

    @staticmethod
    def forward(ctx, f0, f1, w, seed):
        f0, f1 = f0.contiguous(), f1.contiguous()
        N, H, W, C = f0.shape
        out = torch.zeros(N, device=f0.device, dtype=torch.float32)
        wv = w.detach().reshape(-1).float().contiguous()
        if seed is None:
            check(_L().vqb_lpips_tail_fwd(ptr(f0), ptr(f1), ptr(wv), ptr(out), N, H * W, C, stream_ptr()),
                  "lpips_tail_fwd")
        else:
            check(_L().vqb_lpips_tail_fwd_dropout(ptr(f0), ptr(f1), ptr(wv), ptr(out), N, H * W, C, seed, stream_ptr()),
                  "lpips_tail_fwd_dropout")
        ctx.save_for_backward(f0, f1, wv)
        ctx.seed = seed
        return out

    @staticmethod
    def backward(ctx, g):
        f0, f1, wv = ctx.saved_tensors
        N, H, W, C = f0.shape
        g = g.float().contiguous()
        df0 = torch.empty_like(f0)
        if ctx.seed is None:
            check(_L().vqb_lpips_tail_bwd(ptr(f0), ptr(f1), ptr(wv), ptr(g), ptr(df0), N, H * W, C, stream_ptr()),
                  "lpips_tail_bwd")
        else:
            check(_L().vqb_lpips_tail_bwd_dropout(ptr(f0), ptr(f1), ptr(wv), ptr(g), ptr(df0), N, H * W, C, ctx.seed,
                                                  stream_ptr()), "lpips_tail_bwd_dropout")
        return df0, None, None, None


def lpips_tail(f0, f1, w, dropout_seed=None):
    return LpipsTailFn.apply(f0, f1, w, dropout_seed)


def lpips_dropout_mask(seed: int, N: int, HW: int, C: int, device) -> torch.Tensor:
    """The keep mask ([N, HW, C] uint8) the train-mode LPIPS tail kernels use for `seed` (parity tests)."""
    m = torch.empty(N, HW, C, device=device, dtype=torch.uint8)
    check(_L().vqb_lpips_dropout_mask(seed, N, HW, C, ptr(m), stream_ptr()), "lpips_dropout_mask")
    return m


def vq_argmin(z_flat: torch.Tensor, codebook: torch.Tensor):
    """z_flat [M, D] fp32, codebook [K, D] fp32 -> (idx int64 [M], zq fp32 [M, D], sum of squared errors (0-dim))."""
    require_cuda(z_flat)
    z_flat = z_flat.detach().float().contiguous()
    cb = codebook.detach().float().contiguous()
    M, D = z_flat.shape
    idx = torch.empty(M, device=z_flat.device, dtype=torch.int64)
    zq = torch.empty_like(z_flat)
    sq = torch.zeros((), device=z_flat.device, dtype=torch.float32)
    check(_L().vqb_vq_argmin(ptr(z_flat), ptr(cb), ptr(idx), ptr(zq), ptr(sq), M, cb.shape[0], D, stream_ptr()),
          "vq_argmin")
    return idx, zq, sq
