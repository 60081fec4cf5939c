"""Canonical VQ codebook oracle (BASELINE.json config 4). TEST INFRASTRUCTURE (see oracle/__init__.py).

PARITY UNPINNED BY THE REFERENCE: cloneofsimo/vqgan-training contains no codebook / argmin / commitment loss
anywhere, so there is nothing of the reference's to check against. This file pins the standard
VQ-GAN VectorQuantizer semantics instead, with a canonical float32 distance so the index is bit-reproducible:

    d[i][j] = sum_{c = 0..D-1, in order} fma-free float32  (z[i][c] - e[j][c])^2   accumulated left to right
    idx[i]  = smallest j attaining min_j d[i][j]            (torch.argmin tie rule)
    z_q     = e[idx];  loss = beta * mean((sg[z_q] - z)^2) + mean((z_q - sg[z])^2);  out = z + sg[z_q - z]
"""
from __future__ import annotations

import numpy as np


def vq_distances(z: np.ndarray, e: np.ndarray) -> np.ndarray:
    """z [M,D] float32, e [K,D] float32 -> d [M,K] float32 with the canonical accumulation order."""
    z = np.ascontiguousarray(z, dtype=np.float32)
    e = np.ascontiguousarray(e, dtype=np.float32)
    d = np.zeros((z.shape[0], e.shape[0]), dtype=np.float32)
    for c in range(z.shape[1]):
        diff = (z[:, c:c + 1] - e[None, :, c]).astype(np.float32)
        d = (d + (diff * diff).astype(np.float32)).astype(np.float32)
    return d


def vq_argmin(z: np.ndarray, e: np.ndarray) -> np.ndarray:
    return np.argmin(vq_distances(z, e), axis=1).astype(np.int64)  # np.argmin returns the first minimum


def vq_forward(z: np.ndarray, e: np.ndarray, beta: float = 0.25):
    """z [M,D] -> (z_q [M,D], idx [M], loss scalar float32, top2_gap [M])."""
    d = vq_distances(z, e)
    idx = np.argmin(d, axis=1)
    zq = e[idx]
    diff2 = np.mean((zq.astype(np.float64) - z.astype(np.float64)) ** 2)
    loss = np.float32((1.0 + beta) * diff2)
    part = np.partition(d, 1, axis=1)
    return zq, idx.astype(np.int64), loss, (part[:, 1] - part[:, 0])
