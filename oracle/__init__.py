"""oracle/ - CPU restatement of the reference algorithms for the CRB-active-3Ddet hot path.

TEST INFRASTRUCTURE ONLY. Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs
may import this package, and only as the checker or as the timed CPU baseline. The product path
(crb-active-3ddet_b200/) never imports it and has no CPU fallback.

Pinning status (see DESIGN.md "Oracle"):
  * iou3d_nms / roiaware_pool3d / pointnet2_stack: pinned - the reference's own CUDA kernels are compiled from
    /root/reference into oracle/_ref/libpcdet_ref_kernels.so (oracle/build.py) and compared on the GPU box; the C
    restatement in oracle/csrc/oracle.c is validated against golden vectors produced from it (tests/golden/).
  * CRB stage 1/3: pinned - the oracle calls the very library functions the reference calls
    (torch.distributions.Categorical, sklearn KernelDensity, scipy.stats.entropy / uniform).
  * spconv (voxelizer, rulebook, sparse conv): PARITY UNPINNED at the library boundary - spconv-cu113==2.1.21 is a
    third-party dependency absent from /root/reference and from this image. The restatement follows spconv's published
    semantics and is anchored by an independent dense torch.nn.functional.conv3d cross-check.
"""
import ctypes
import os

import numpy as np

from . import build as _build

_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(_build.build_oracle())
        _lib.oracle_box_overlap.restype = ctypes.c_float
        _lib.oracle_iou_bev.restype = ctypes.c_float
    return _lib


def ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def i32(a):
    return np.ascontiguousarray(a, dtype=np.int32)
