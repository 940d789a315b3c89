"""Shared helpers for the parity tests."""
import ctypes
import os

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
_ref = None


def ref_kernels():
    """The reference's own CUDA kernels (oracle/_ref, built from /root/reference by oracle/build.py). None if absent."""
    global _ref
    if _ref is None:
        from oracle import build
        path = build.build_ref()
        if path is None or not os.path.exists(path):
            return None
        _ref = ctypes.CDLL(path)
    return _ref


_ref_batch = None


def ref_batch_kernels():
    """The reference's pointnet2_batch / voxel_query / roipoint_pool3d kernels (oracle/_ref, second library). None if absent."""
    global _ref_batch
    if _ref_batch is None:
        from oracle import build
        path = build.build_ref_batch()
        if path is None or not os.path.exists(path):
            return None
        _ref_batch = ctypes.CDLL(path)
    return _ref_batch


def P(t):
    return ctypes.c_void_p(t.data_ptr())


def rand_boxes(rng, n, extent=30.0, cluster=False):
    """(n,7) float32 boxes; cluster=True piles them up so that many pairs overlap."""
    if cluster:
        centres = rng.uniform(-extent, extent, (max(n // 12, 1), 2))
        xy = centres[rng.integers(0, len(centres), n)] + rng.normal(0, 0.8, (n, 2))
    else:
        xy = rng.uniform(-extent, extent, (n, 2))
    z = rng.uniform(-1.5, 0.5, (n, 1))
    size = rng.uniform([1.0, 0.5, 1.0], [5.0, 2.5, 2.0], (n, 3))
    yaw = rng.uniform(-np.pi, np.pi, (n, 1))
    return np.concatenate([xy, z, size, yaw], 1).astype(np.float32)


def cu(a, dev, dtype=None):
    t = torch.as_tensor(np.ascontiguousarray(a))
    if dtype is not None:
        t = t.to(dtype)
    return t.to(dev)
