"""Szudzik pairing of (time bucket, pad) -- host-side mirror of `detector/pairing.py:5-55`.

The device uses the same integer arithmetic (`csrc/attpc_kernels.cuh: szudzik_pair`); keys
stay below 2^27 (pad <= 10239, tb <= 10239), int32 on device and int64 at the boundary.
Works on Python ints and numpy integer arrays alike.
"""

from __future__ import annotations

import numpy as np


def pair(tb, pad):
    """``tb*tb + tb + pad`` if ``tb >= pad`` else ``pad*pad + tb``; -1 for negative input."""
    if np.isscalar(tb) and np.isscalar(pad):
        if tb < 0 or pad < 0:
            return -1
        return tb * tb + tb + pad if tb >= pad else pad * pad + tb
    tb = np.asarray(tb, dtype=np.int64)
    pad = np.asarray(pad, dtype=np.int64)
    key = np.where(tb >= pad, tb * tb + tb + pad, pad * pad + tb)
    return np.where((tb < 0) | (pad < 0), -1, key)


def unpair(id):
    """Inverse of :func:`pair`; returns ``(tb, pad)`` (floats for scalars, like the reference)."""
    if np.isscalar(id):
        if id < 0:
            return (-1, -1)
        s = float(np.floor(np.sqrt(id)))
        if id - s * s < s:
            return (id - s * s, s)
        return (s, id - s * s - s)
    key = np.asarray(id, dtype=np.int64)
    s = np.floor(np.sqrt(key.astype(np.float64))).astype(np.int64)
    # guard the float sqrt against off-by-one at perfect squares
    s = np.where(s * s > key, s - 1, s)
    s = np.where((s + 1) * (s + 1) <= key, s + 1, s)
    rem = key - s * s
    tb = np.where(rem < s, rem, s)
    pad = np.where(rem < s, s, rem - s)
    neg = key < 0
    return np.where(neg, -1, tb), np.where(neg, -1, pad)
