"""PointNet++ stacked-op oracle (TEST INFRASTRUCTURE ONLY); C code in csrc/oracle.c.

Follows pcdet/ops/pointnet2/pointnet2_stack/src/{ball_query_gpu.cu:16-66, group_points_gpu.cu:15-102,
sampling_gpu.cu:9-140, interpolate_gpu.cu:16-126} and pointnet2_utils.py:8-184 for the calling conventions.
"""
import math

import numpy as np

from . import f32, i32, lib, ptr


def ball_query(radius, nsample, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt):
    xyz, new_xyz = f32(xyz), f32(new_xyz)
    xc, nc = i32(xyz_batch_cnt), i32(new_xyz_batch_cnt)
    idx = np.zeros((len(new_xyz), nsample), np.int32)
    import ctypes
    lib().oracle_ball_query(len(xc), ctypes.c_float(radius), int(nsample), ptr(new_xyz), ptr(nc), ptr(xyz), ptr(xc), ptr(idx))
    return idx


def group_points(features, features_batch_cnt, idx, idx_batch_cnt):
    features, idx = f32(features), i32(idx)
    fc, ic = i32(features_batch_cnt), i32(idx_batch_cnt)
    M, ns = idx.shape
    C = features.shape[1]
    out = np.zeros((M, C, ns), np.float32)
    lib().oracle_group_points(len(fc), C, ns, ptr(features), ptr(fc), ptr(idx), ptr(ic), ptr(out))
    return out


def group_points_grad(grad_out, idx, idx_batch_cnt, features_batch_cnt, n):
    grad_out, idx = f32(grad_out), i32(idx)
    M, C, ns = grad_out.shape
    g = np.zeros((n, C), np.float64)
    m0 = f0 = 0
    for b in range(len(idx_batch_cnt)):
        for m in range(m0, m0 + int(idx_batch_cnt[b])):
            np.add.at(g, f0 + idx[m], grad_out[m].T.astype(np.float64))
        m0 += int(idx_batch_cnt[b]); f0 += int(features_batch_cnt[b])
    return g.astype(np.float32)


def ref_block(n):
    """Thread count the reference launches FPS with (sampling_gpu.cu:9-13)."""
    pow_2 = int(math.log(float(n)) / math.log(2.0))
    return max(min(1 << pow_2, 1024), 1)


def farthest_point_sampling(points, m, block=None):
    """points (n,3) one batch element -> (idx (m,), final temp (n,))."""
    points = f32(points)
    n = len(points)
    temp = np.full((n,), 1e10, np.float32)
    idx = np.zeros((m,), np.int32)
    lib().oracle_fps(n, int(m), int(block or ref_block(n)), ptr(points), ptr(temp), ptr(idx))
    return idx, temp


def three_nn(unknown, unknown_batch_cnt, known, known_batch_cnt):
    unknown, known = f32(unknown), f32(known)
    uc, kc = i32(unknown_batch_cnt), i32(known_batch_cnt)
    d2 = np.zeros((len(unknown), 3), np.float32)
    idx = np.zeros((len(unknown), 3), np.int32)
    lib().oracle_three_nn(len(uc), ptr(unknown), ptr(uc), ptr(known), ptr(kc), ptr(d2), ptr(idx))
    return d2, idx


def three_interpolate(features, idx, weight):
    features, weight = f32(features), f32(weight)
    return (weight[:, 0:1] * features[idx[:, 0]] + weight[:, 1:2] * features[idx[:, 1]] + weight[:, 2:3] * features[idx[:, 2]]).astype(np.float32)


def voxel_query(max_range, radius, nsample, xyz, new_xyz, new_coords, point_indices):
    """pointnet2_stack/voxel_query_utils.py:12-42 + src/voxel_query_gpu.cu:13-98 -> raw idx (M, nsample) (idx[m, 0] = -1: empty)."""
    import ctypes
    xyz, new_xyz = f32(xyz), f32(new_xyz)
    new_coords, point_indices = i32(new_coords), i32(point_indices)
    B, Z, Y, X = point_indices.shape
    M = len(new_xyz)
    idx = np.zeros((M, nsample), np.int32)
    zr, yr, xr = max_range
    lib().oracle_voxel_query(M, Z, Y, X, int(nsample), ctypes.c_float(radius), int(zr), int(yr), int(xr), ptr(new_xyz), ptr(xyz),
                             ptr(new_coords), ptr(point_indices), ptr(idx))
    return idx


def roipoint_pool3d(xyz, boxes, feat, S):
    """roipoint_pool3d_kernel.cu:38-165 for one frame: xyz (N,3), boxes (M,7) (already enlarged), feat (N,C) ->
    pooled (M,S,3+C), empty (M,) int32."""
    xyz, boxes, feat = f32(xyz), f32(boxes), f32(feat)
    N, M, C = len(xyz), len(boxes), feat.shape[1]
    pooled = np.zeros((M, S, 3 + C), np.float32)
    empty = np.zeros((M,), np.int32)
    lib().oracle_roipoint_pool3d(N, M, C, int(S), ptr(xyz), ptr(boxes), ptr(feat), ptr(pooled), ptr(empty))
    return pooled, empty
