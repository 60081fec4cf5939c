"""Functional fp32 CPU restatement of utils.LPIPS / vgg16 / PatchDiscriminator / wavelet (reference utils.py).
TEST INFRASTRUCTURE (see oracle/__init__.py). Pinned by tests/golden/lpips_*.npz, patchd_*.npz, wavelet_*.npz.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

# torchvision vgg16.features conv indices per slice (utils.py:102-111,150-154); a max-pool precedes slices 2..5
VGG_SLICES = [[0, 2], [5, 7], [10, 12, 14], [17, 19, 21], [24, 26, 28]]
VGG_CHANNELS = [(3, 64), (64, 64), (64, 128), (128, 128), (128, 256), (256, 256), (256, 256), (256, 512), (512, 512),
                (512, 512), (512, 512), (512, 512), (512, 512)]
LPIPS_CHNS = [64, 128, 256, 512, 512]  # utils.py:13
SHIFT = torch.tensor([-0.030, -0.088, -0.188])  # utils.py:63-68
SCALE = torch.tensor([0.458, 0.448, 0.450])


def scaling_layer(x):  # utils.py:63-71: shift/scale are fp32 buffers -> a bf16 input (autocast decoder output) promotes to fp32
    return (x - SHIFT.to(x.device)[None, :, None, None]) / SCALE.to(x.device)[None, :, None, None]


def vgg_features(sd, x, key_fmt):
    """key_fmt(slice_idx(1-based), conv_idx) -> state_dict key prefix of that conv.
    Returns [relu1_2, relu2_2, relu3_3, relu4_3, relu5_3] (utils.py:113-131)."""
    outs = []
    h = x
    for s, idxs in enumerate(VGG_SLICES):
        if s > 0:
            h = F.max_pool2d(h, 2)
        for ci in idxs:
            p = key_fmt(s + 1, ci)
            h = F.relu(F.conv2d(h, sd[p + ".weight"], sd[p + ".bias"], padding=1))
        outs.append(h)
    return outs


def lpips_key(s, ci):  # LPIPS: net.slice{s}.{idx}  (utils.py:92-111)
    return f"net.slice{s}.{ci}"


def patchd_key(s, ci):
    # PatchDiscriminator: nn.Sequential(_vgg.features[a:b]) (utils.py:150-154); slicing a Sequential keeps the
    # ORIGINAL module names, so the key is slice{s}.0.{torchvision idx}
    return f"slice{s}.0.{ci}"


def normalize_tensor(x, eps=1e-10):  # utils.py:134-136 (eps added AFTER the sqrt)
    return x / (torch.sqrt(torch.sum(x ** 2, dim=1, keepdim=True)) + eps)


def lpips_forward(sd, inp, tgt, keep_masks=None):
    """utils.py:39-57 -> [B,1,1,1]. keep_masks=None: eval mode (dropout off). keep_masks = five [B,C,H,W] 0/1 tensors:
    the train-mode nn.Dropout(0.5) of NetLinLayer (utils.py:79-89) with that explicit keep mask, i.e. exactly
    F.dropout's arithmetic `x * mask / (1 - p)` in front of the 1x1 lin conv."""
    f0 = vgg_features(sd, scaling_layer(inp), lpips_key)
    f1 = vgg_features(sd, scaling_layer(tgt), lpips_key)
    val = None
    for kk in range(5):
        d = (normalize_tensor(f0[kk]) - normalize_tensor(f1[kk])) ** 2
        if keep_masks is not None:
            d = d * keep_masks[kk].to(d.dtype) * 2.0
        r = F.conv2d(d, sd[f"lin{kk}.model.1.weight"]).mean([2, 3], keepdim=True)  # utils.py:51-53,139-140
        val = r if val is None else val + r
    return val


def patchd_forward(sd, x):
    """utils.py:187-203 -> [B, (H/16)(W/16)]."""
    f = vgg_features(sd, scaling_layer(x), patchd_key)

    def c(p, t, stride):
        return F.conv2d(t, sd[p + ".weight"], sd[p + ".bias"], stride=stride)

    bc1 = c("binary_classifier1.2", F.relu(c("binary_classifier1.0", f[0], 4)), 4).flatten(1)
    bc2 = c("binary_classifier2.2", F.relu(c("binary_classifier2.0", f[1], 4)), 2).flatten(1)
    bc3 = c("binary_classifier3.2", F.relu(c("binary_classifier3.0", f[2], 2)), 2).flatten(1)
    bc4 = c("binary_classifier4.0", f[3], 2).flatten(1)
    bc5 = c("binary_classifier5.0", f[4], 1).flatten(1)
    return bc1 + bc2 + bc3 + bc4 + bc5


def lpips_state_dict_shapes() -> dict:
    sh = {"scaling_layer.shift": (1, 3, 1, 1), "scaling_layer.scale": (1, 3, 1, 1)}
    n = 0
    for s, idxs in enumerate(VGG_SLICES):
        for ci in idxs:
            cin, cout = VGG_CHANNELS[n]
            n += 1
            sh[f"net.slice{s + 1}.{ci}.weight"] = (cout, cin, 3, 3)
            sh[f"net.slice{s + 1}.{ci}.bias"] = (cout,)
    for kk, c in enumerate(LPIPS_CHNS):
        sh[f"lin{kk}.model.1.weight"] = (1, c, 1, 1)
    return sh


def patchd_state_dict_shapes() -> dict:
    sh = {"scaling_layer.shift": (1, 3, 1, 1), "scaling_layer.scale": (1, 3, 1, 1)}
    n = 0
    for s, idxs in enumerate(VGG_SLICES):
        for ci in idxs:
            cin, cout = VGG_CHANNELS[n]
            n += 1
            p = patchd_key(s + 1, ci)
            sh[p + ".weight"] = (cout, cin, 3, 3)
            sh[p + ".bias"] = (cout,)
    heads = {"binary_classifier1.0": (32, 64, 4), "binary_classifier1.2": (1, 32, 4),
             "binary_classifier2.0": (64, 128, 4), "binary_classifier2.2": (1, 64, 2),
             "binary_classifier3.0": (128, 256, 2), "binary_classifier3.2": (1, 128, 2),
             "binary_classifier4.0": (1, 512, 2), "binary_classifier5.0": (1, 512, 1)}
    for p, (co, ci, k) in heads.items():
        sh[p + ".weight"] = (co, ci, k, k)
        sh[p + ".bias"] = (co,)
    return sh


# ---- wavelet front-end (utils.py:206-247) -------------------------------------------------------
_DEC_LO = torch.tensor([-0.1768, 0.3536, 1.0607, 0.3536, -0.1768, 0.0000])
_DEC_HI = torch.tensor([0.0000, -0.0000, 0.3536, -0.7071, 0.3536, -0.0000])


def wavelet_filters():  # utils.py:211-221  -> [4,1,6,6]
    f = torch.stack([_DEC_LO[None, :] * _DEC_LO[:, None], _DEC_LO[None, :] * _DEC_HI[:, None],
                     _DEC_HI[None, :] * _DEC_LO[:, None], _DEC_HI[None, :] * _DEC_HI[:, None]], dim=0)
    return f.unsqueeze(1)


def wavelet_transform_multi_channel(x):  # utils.py:229-247 -> [B,4C,H/2,W/2], channel order (c, band)
    B, C, H, W = x.shape
    padded = F.pad(x, (2, 2, 2, 2))
    w = wavelet_filters().to(x)
    res = torch.cat([F.conv2d(padded[:, c:c + 1], w, stride=2) for c in range(C)], dim=1)
    return res
