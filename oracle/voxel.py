"""Voxelizer + MeanVFE oracle (TEST INFRASTRUCTURE ONLY).

Restates spconv 2.1 `Point2VoxelCPU3d.point_to_voxel` as called from
pcdet/datasets/processor/data_processor.py:15-60,115-143 and pcdet/models/backbones_3d/vfe/mean_vfe.py:14-31.
spconv is not vendored in the reference (PARITY UNPINNED at that boundary); semantics per SURVEY.md 2.4.
"""
import ctypes

import numpy as np

from . import f32, i32, lib, ptr


def grid_size(pc_range, voxel_size):
    pc_range = np.asarray(pc_range, dtype=np.float64)
    voxel_size = np.asarray(voxel_size, dtype=np.float64)
    return np.round((pc_range[3:6] - pc_range[0:3]) / voxel_size).astype(np.int64)


def point_to_voxel(points, pc_range, voxel_size, max_pts, max_voxels, n_feat=None):
    """points (n, C) f32 of ONE frame -> (voxels (M,P,C), coords (M,3) zyx, num (M)) (C implementation)."""
    points = f32(points)
    n, stride = points.shape
    n_feat = stride if n_feat is None else n_feat
    grid = i32(grid_size(pc_range, voxel_size))
    voxels = np.zeros((max_voxels, max_pts, n_feat), dtype=np.float32)
    coords = np.zeros((max_voxels, 3), dtype=np.int32)
    num = np.zeros((max_voxels,), dtype=np.int32)
    m = lib().oracle_voxelize(ptr(points), n, stride, n_feat, ptr(f32(pc_range)), ptr(f32(voxel_size)), ptr(grid),
                              int(max_pts), int(max_voxels), ptr(voxels), ptr(coords), ptr(num))
    return voxels[:m].copy(), coords[:m].copy(), num[:m].copy()


def point_to_voxel_py(points, pc_range, voxel_size, max_pts, max_voxels):
    """Pure-Python statement of the same serial algorithm (small inputs; cross-checks the C version)."""
    points = f32(points)
    lo = f32(pc_range)[:3]
    vs = f32(voxel_size)
    grid = grid_size(pc_range, voxel_size)
    table = {}
    voxels, coords, num = [], [], []
    for p in points:
        c = np.floor((p[:3] - lo) / vs)  # float32 arithmetic
        if np.any(c < 0) or np.any(c >= grid.astype(np.float32)) or np.any(np.isnan(c)):
            continue
        key = (int(c[2]), int(c[1]), int(c[0]))
        vid = table.get(key)
        if vid is None:
            if len(voxels) >= max_voxels:
                continue
            vid = len(voxels)
            table[key] = vid
            voxels.append(np.zeros((max_pts, points.shape[1]), dtype=np.float32))
            coords.append(key)
            num.append(0)
        if num[vid] < max_pts:
            voxels[vid][num[vid]] = p
            num[vid] += 1
    if not voxels:
        return (np.zeros((0, max_pts, points.shape[1]), np.float32), np.zeros((0, 3), np.int32), np.zeros((0,), np.int32))
    return np.stack(voxels), np.asarray(coords, dtype=np.int32), np.asarray(num, dtype=np.int32)


def mean_vfe(voxels, num):
    voxels = f32(voxels)
    num = i32(num)
    m, p, c = voxels.shape
    out = np.zeros((m, c), dtype=np.float32)
    if m:
        lib().oracle_mean_vfe(ptr(voxels), ptr(num), m, p, c, ptr(out))
    return out


def voxelize_batch(frames, pc_range, voxel_size, max_pts, max_voxels):
    """List of per-frame point arrays -> collated (mean feats (M,C), coords (M,4) [b,z,y,x], num (M), frame offsets),
    mirroring DatasetTemplate.collate_batch (pcdet/datasets/dataset.py:173-178) + MeanVFE."""
    feats, coords, nums, offs = [], [], [], [0]
    for b, pts in enumerate(frames):
        v, c, n = point_to_voxel(pts, pc_range, voxel_size, max_pts, max_voxels)
        feats.append(mean_vfe(v, n))
        coords.append(np.concatenate([np.full((len(c), 1), b, np.int32), c], axis=1))
        nums.append(n)
        offs.append(offs[-1] + len(c))
    return np.concatenate(feats), np.concatenate(coords), np.concatenate(nums), np.asarray(offs, np.int32)
