"""Rotated IoU / NMS / points-in-boxes / RoI-aware pooling oracle (TEST INFRASTRUCTURE ONLY); C code in csrc/oracle.c.

Follows pcdet/ops/iou3d_nms/src/iou3d_cpu.cpp:60-252, iou3d_nms.cpp:90-136, iou3d_nms_kernel.cu:314-325,
pcdet/ops/roiaware_pool3d/src/roiaware_pool3d_kernel.cu:16-190,313-336, roiaware_pool3d.cpp:119-168,
pcdet/ops/iou3d_nms/iou3d_nms_utils.py:48-81 (boxes_iou3d), pcdet/models/detectors/detector3d_template.py:379-387.
"""
import numpy as np

from . import f32, i32, lib, ptr


def boxes_overlap_bev(a, b):
    a, b = f32(a)[:, :7].copy(), f32(b)[:, :7].copy()
    out = np.zeros((len(a), len(b)), np.float32)
    lib().oracle_pairwise(ptr(a), len(a), ptr(b), len(b), 0, ptr(out))
    return out


def boxes_iou_bev(a, b):
    a, b = f32(a)[:, :7].copy(), f32(b)[:, :7].copy()
    out = np.zeros((len(a), len(b)), np.float32)
    lib().oracle_pairwise(ptr(a), len(a), ptr(b), len(b), 1, ptr(out))
    return out


def boxes_iou3d(a, b):
    a, b = f32(a), f32(b)
    ov = boxes_overlap_bev(a, b)
    amax, amin = (a[:, 2] + a[:, 5] / 2)[:, None], (a[:, 2] - a[:, 5] / 2)[:, None]
    bmax, bmin = (b[:, 2] + b[:, 5] / 2)[None, :], (b[:, 2] - b[:, 5] / 2)[None, :]
    oh = np.clip(np.minimum(amax, bmax) - np.maximum(amin, bmin), 0, None)
    o3 = ov * oh
    va, vb = (a[:, 3] * a[:, 4] * a[:, 5])[:, None], (b[:, 3] * b[:, 4] * b[:, 5])[None, :]
    return (o3 / np.clip(va + vb - o3, 1e-6, None)).astype(np.float32)


def nms_sorted(boxes_sorted, thresh, rotated=True, return_iou=False):
    """Greedy NMS over boxes sorted by descending score -> kept indices (ascending)."""
    b = f32(boxes_sorted)[:, :7].copy()
    n = len(b)
    keep = np.zeros((max(n, 1),), np.int64)
    iou = np.zeros((n, n), np.float32) if return_iou else None
    nk = lib().oracle_nms(ptr(b), n, _cf(thresh),
                          int(bool(rotated)), ptr(keep), ptr(iou) if iou is not None else None)
    return (keep[:nk].copy(), iou) if return_iou else keep[:nk].copy()


def _cf(x):
    import ctypes
    return ctypes.c_float(float(x))


def points_in_boxes(boxes, pts):
    """One frame: boxes (T,7), pts (M,>=3) -> (M,) first containing box or -1 (GPU predicate, MARGIN 1e-5)."""
    boxes, pts = f32(boxes)[:, :7].copy(), f32(pts)
    out = np.full((len(pts),), -1, np.int32)
    if len(pts):
        lib().oracle_points_in_boxes(ptr(boxes), len(boxes), ptr(pts), len(pts), pts.shape[1], ptr(out))
    return out


def points_in_boxes_cpu(boxes, pts):
    boxes, pts = f32(boxes)[:, :7].copy(), f32(pts)[:, :3].copy()
    out = np.zeros((len(boxes), len(pts)), np.int32)
    if len(boxes) and len(pts):
        lib().oracle_points_in_boxes_cpu(ptr(boxes), len(boxes), ptr(pts), len(pts), ptr(out))
    return out


def box_density(boxes, pts):
    """detector3d_template.py:379-387: (#points whose FIRST containing box is j) / (dx*dy*dz), float32."""
    boxes = f32(boxes)
    idx = points_in_boxes(boxes, pts)
    cnt = np.bincount(idx[idx >= 0], minlength=len(boxes)).astype(np.float32)
    vol = (boxes[:, 3] * boxes[:, 4]) * boxes[:, 5]
    return cnt / vol, cnt.astype(np.int32), idx


def roiaware_pool3d(rois, pts, feat, out_size, max_pts_each_voxel, method):
    rois, pts, feat = f32(rois)[:, :7].copy(), f32(pts)[:, :3].copy(), f32(feat)
    ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else out_size
    n, C = len(rois), feat.shape[1]
    pooled = np.zeros((n, ox, oy, oz, C), np.float32)
    argmax = np.zeros((n, ox, oy, oz, C), np.int32)
    pidx = np.zeros((n, ox, oy, oz, max_pts_each_voxel), np.int32)
    lib().oracle_roiaware_pool(ptr(rois), n, ptr(pts), ptr(feat), len(pts), C, max_pts_each_voxel, ox, oy, oz,
                               {"max": 0, "avg": 1}[method], ptr(argmax), ptr(pidx), ptr(pooled))
    return pooled, argmax, pidx
