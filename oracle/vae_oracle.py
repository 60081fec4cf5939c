"""Functional fp32 CPU restatement of the reference VAE (ae.py). TEST INFRASTRUCTURE (see oracle/__init__.py).

All functions take a reference-format state_dict `sd` (keys as produced by ae.VAE(...).state_dict(), fp32 OIHW)
and a key prefix. Pinned against the imported reference by tests/golden/vae_*.npz.
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import List

import torch
import torch.nn.functional as F


@dataclass
class VAEConfig:
    resolution: int = 256
    in_channels: int = 3
    ch: int = 128
    out_ch: int = 3
    ch_mult: tuple = (1, 2, 4, 4)
    num_res_blocks: int = 2
    z_channels: int = 16
    use_attn: bool = False
    decoder_also_perform_hr: bool = False

    @property
    def dec_ch_mult(self):
        # ae.py:381  ch_mult + [4] when the decoder also performs 2x "HR" up-sampling
        return tuple(self.ch_mult) + ((4,) if self.decoder_also_perform_hr else ())


def swish(x):  # ae.py:13-14
    return x * torch.sigmoid(x)


def group_norm(sd, p, x):  # ae.py:41-53  FP32GroupNorm: 32 groups, eps 1e-6, fp32 math
    return F.group_norm(x.float(), 32, sd[p + ".weight"].float(), sd[p + ".bias"].float(), 1e-6).type_as(x)


def conv(sd, p, x, stride=1, padding=1):
    return F.conv2d(x, sd[p + ".weight"], sd.get(p + ".bias"), stride=stride, padding=padding)


def resnet_block(sd, p, x):  # ae.py:124-140
    h = conv(sd, p + ".conv1", swish(group_norm(sd, p + ".norm1", x)))
    h = conv(sd, p + ".conv2", swish(group_norm(sd, p + ".norm2", h)))
    if (p + ".nin_shortcut.weight") in sd:
        x = conv(sd, p + ".nin_shortcut", x, padding=0)
    return x + h


def attn_block(sd, p, x):  # ae.py:74-93: GN -> 1x1 qkv (no bias) -> heads of 64 -> softmax(qk^T/8)v -> 1x1 proj -> +x
    h = group_norm(sd, p + ".norm", x)
    qkv = F.conv2d(h, sd[p + ".qkv.weight"])
    q, k, v = qkv.chunk(3, dim=1)
    b, c, hh, ww = q.shape
    nh, hd = c // 64, 64

    def split(t):  # "b (h d) x y -> b h (x y) d"
        return t.reshape(b, nh, hd, hh * ww).permute(0, 1, 3, 2)

    q, k, v = split(q), split(k), split(v)
    att = torch.softmax(q @ k.transpose(-1, -2) / math.sqrt(hd), dim=-1) @ v
    att = att.permute(0, 1, 3, 2).reshape(b, c, hh, ww)  # "b h (x y) d -> b (h d) x y"
    return x + F.conv2d(att, sd[p + ".proj_out.weight"])


def downsample(sd, p, x):  # ae.py:150-154: zero pad right/bottom, conv3x3 stride 2 no padding
    return conv(sd, p + ".conv", F.pad(x, (0, 1, 0, 1)), stride=2, padding=0)


def upsample(sd, p, x):  # ae.py:164-167: nearest x2 then conv3x3 p1
    return conv(sd, p + ".conv", F.interpolate(x, scale_factor=2.0, mode="nearest"))


def encoder_forward(sd, x, cfg: VAEConfig, p="encoder"):  # ae.py:239-257 (no wavelet)
    nres = len(cfg.ch_mult)
    h = conv(sd, p + ".conv_in", x)
    for i in range(nres):
        for j in range(cfg.num_res_blocks):
            h = resnet_block(sd, f"{p}.down.{i}.block.{j}", h)
        if i != nres - 1:
            h = downsample(sd, f"{p}.down.{i}.downsample", h)
    h = resnet_block(sd, p + ".mid.block_1", h)
    if (p + ".mid.attn_1.qkv.weight") in sd:
        h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    return conv(sd, p + ".conv_out", swish(group_norm(sd, p + ".norm_out", h)))


def decoder_forward(sd, z, cfg: VAEConfig, p="decoder"):  # ae.py:318-333
    mult = cfg.dec_ch_mult
    nres = len(mult)
    h = conv(sd, p + ".conv_in", z)
    h = resnet_block(sd, p + ".mid.block_1", h)
    if (p + ".mid.attn_1.qkv.weight") in sd:
        h = attn_block(sd, p + ".mid.attn_1", h)
    h = resnet_block(sd, p + ".mid.block_2", h)
    for i in reversed(range(nres)):
        for j in range(cfg.num_res_blocks + 1):
            h = resnet_block(sd, f"{p}.up.{i}.block.{j}", h)
        if i != 0:
            h = upsample(sd, f"{p}.up.{i}.upsample", h)
    return conv(sd, p + ".conv_out", swish(group_norm(sd, p + ".norm_out", h)))


def reg(z):  # ae.py:342-348 DiagonalGaussian with std = 0.00: identity for finite z
    return z


def vae_forward(sd, x, cfg: VAEConfig):  # ae.py:388-392 -> (decz, z)
    z = encoder_forward(sd, x, cfg)
    return decoder_forward(sd, reg(z), cfg), z


def state_dict_shapes(cfg: VAEConfig) -> dict:
    """Key -> shape of ae.VAE(...).state_dict() for cfg (derived from ae.py:170-386); used to build seeded weights
    without importing the reference."""
    sh = {}

    def conv_(p, cin, cout, k, bias=True):
        sh[p + ".weight"] = (cout, cin, k, k)
        if bias:
            sh[p + ".bias"] = (cout,)

    def norm_(p, c):
        sh[p + ".weight"] = (c,)
        sh[p + ".bias"] = (c,)

    def res_(p, cin, cout):
        norm_(p + ".norm1", cin)
        conv_(p + ".conv1", cin, cout, 3)
        norm_(p + ".norm2", cout)
        conv_(p + ".conv2", cout, cout, 3)
        if cin != cout:
            conv_(p + ".nin_shortcut", cin, cout, 1)

    def attn_(p, c):
        norm_(p + ".norm", c)
        conv_(p + ".qkv", c, 3 * c, 1, bias=False)
        conv_(p + ".proj_out", c, c, 1, bias=False)

    ch, mult = cfg.ch, tuple(cfg.ch_mult)
    # encoder
    conv_("encoder.conv_in", cfg.in_channels, ch, 3)
    in_mult = (1,) + mult
    block_in = ch
    for i in range(len(mult)):
        block_in, block_out = ch * in_mult[i], ch * mult[i]
        for j in range(cfg.num_res_blocks):
            res_(f"encoder.down.{i}.block.{j}", block_in, block_out)
            block_in = block_out
        if i != len(mult) - 1:
            conv_(f"encoder.down.{i}.downsample.conv", block_in, block_in, 3)
    res_("encoder.mid.block_1", block_in, block_in)
    if cfg.use_attn:
        attn_("encoder.mid.attn_1", block_in)
    res_("encoder.mid.block_2", block_in, block_in)
    norm_("encoder.norm_out", block_in)
    conv_("encoder.conv_out", block_in, cfg.z_channels, 3)
    # decoder
    dm = cfg.dec_ch_mult
    block_in = ch * dm[-1]
    conv_("decoder.conv_in", cfg.z_channels, block_in, 3)
    res_("decoder.mid.block_1", block_in, block_in)
    if cfg.use_attn:
        attn_("decoder.mid.attn_1", block_in)
    res_("decoder.mid.block_2", block_in, block_in)
    for i in reversed(range(len(dm))):
        block_out = ch * dm[i]
        for j in range(cfg.num_res_blocks + 1):
            res_(f"decoder.up.{i}.block.{j}", block_in, block_out)
            block_in = block_out
        if i != 0:
            conv_(f"decoder.up.{i}.upsample.conv", block_in, block_in, 3)
    norm_("decoder.norm_out", block_in)
    conv_("decoder.conv_out", block_in, cfg.out_ch, 3)
    return sh
