"""Deterministic, torch-RNG-independent tensors so fixtures need to store outputs only.

`tensor(name, shape, scale)` draws from numpy's PCG64 seeded by a stable hash of `name`; the same call
returns bit-identical values in make_golden.py (build container, with the reference imported) and in
the tests (anywhere).
"""
from __future__ import annotations

import hashlib

import numpy as np
import torch


def _seed(name: str) -> int:
    return int.from_bytes(hashlib.sha256(name.encode()).digest()[:8], "little")


def tensor(name: str, shape, scale: float = 1.0, kind: str = "normal") -> torch.Tensor:
    rng = np.random.Generator(np.random.PCG64(_seed(name)))
    if kind == "normal":
        a = rng.standard_normal(size=tuple(shape), dtype=np.float32) * np.float32(scale)
    elif kind == "uniform":  # [-scale, scale)
        a = (rng.random(size=tuple(shape), dtype=np.float32) * 2 - 1) * np.float32(scale)
    elif kind == "positive":  # [0, scale)
        a = rng.random(size=tuple(shape), dtype=np.float32) * np.float32(scale)
    else:
        raise ValueError(kind)
    return torch.from_numpy(np.ascontiguousarray(a))


def fill_state_dict(sd: dict, tag: str, gain: float = 1.0) -> dict:
    """Seeded values for every floating tensor of a reference-format state_dict (shapes taken from `sd`).

    conv / linear weights ~ N(0, gain^2/fan_in) (keeps activations O(1) through deep stacks),
    norm weights ~ 1 + 0.1 N, biases ~ 0.05 N, LPIPS lin weights non-negative, buffers untouched."""
    out = {}
    for k, v in sd.items():
        if not torch.is_floating_point(v) or k.endswith("scaling_layer.shift") or k.endswith("scaling_layer.scale"):
            out[k] = v.clone()
            continue
        name = f"{tag}/{k}"
        if v.ndim == 4:
            fan_in = v.shape[1] * v.shape[2] * v.shape[3]
            if ".lin" in k or k.startswith("lin"):
                out[k] = tensor(name, v.shape, 1.0 / fan_in, "positive")
            else:
                out[k] = tensor(name, v.shape, gain * (2.0 / fan_in) ** 0.5)
        elif v.ndim == 1 and ("norm" in k and k.endswith("weight")):
            out[k] = 1.0 + tensor(name, v.shape, 0.1)
        elif v.ndim == 1:
            out[k] = tensor(name, v.shape, 0.05)
        else:
            out[k] = tensor(name, v.shape, 0.1)
    return out
