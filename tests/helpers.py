"""Shared helpers for the test-suite: golden loading, seeded state_dicts, error metrics."""
import os

import numpy as np
import torch

from oracle import seeded

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def golden(name):
    d = np.load(os.path.join(GOLDEN, name + ".npz"), allow_pickle=False)
    return {k: d[k] for k in d.files}


def t(a):
    return torch.from_numpy(np.asarray(a))


def rel_l2(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu(), torch.as_tensor(b).detach().double().cpu()
    return ((a - b).norm() / (b.norm() + 1e-30)).item()


def cosine(a, b):
    a, b = torch.as_tensor(a).detach().double().cpu().flatten(), torch.as_tensor(b).detach().double().cpu().flatten()
    return (a @ b / (a.norm() * b.norm() + 1e-30)).item()


def seeded_sd(shapes: dict, tag: str):
    """Same values as oracle/make_golden.py put into the reference modules (seeded.fill_state_dict on the shape table)."""
    from oracle import lpips_oracle as LP

    proto = {}
    for k, shp in shapes.items():
        if k.endswith("scaling_layer.shift"):
            proto[k] = LP.SHIFT.clone().reshape(1, 3, 1, 1)
        elif k.endswith("scaling_layer.scale"):
            proto[k] = LP.SCALE.clone().reshape(1, 3, 1, 1)
        else:
            proto[k] = torch.zeros(shp)
    return seeded.fill_state_dict(proto, tag)
