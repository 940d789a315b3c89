"""spconv.utils drop-in. Only `Point2VoxelCPU3d` is exported: VoxelGeneratorWrapper probes VoxelGeneratorV2, then
VoxelGenerator, then Point2VoxelCPU3d (pcdet/datasets/processor/data_processor.py:17-26) and must land on spconv_ver=2.
The voxelizer runs on the host (DataLoader workers have no CUDA context) through the C ABI's host entry point; the
device voxelizer used by the scoring path is crb3d.ops.voxelize / pcdet_ops.voxel.hard_voxelize."""
import ctypes

import numpy as np

from crb3d import _lib
from cumm import tensorview as tv


class Point2VoxelCPU3d(object):
    def __init__(self, vsize_xyz, coors_range_xyz, num_point_features, max_num_voxels, max_num_points_per_voxel):
        self.vsize = np.asarray(vsize_xyz, dtype=np.float32)
        self.coors_range = np.asarray(coors_range_xyz, dtype=np.float32)
        self.num_point_features = int(num_point_features)
        self.max_num_voxels = int(max_num_voxels)
        self.max_num_points_per_voxel = int(max_num_points_per_voxel)
        rng = np.asarray(coors_range_xyz, dtype=np.float64)
        self.grid_size = np.round((rng[3:6] - rng[0:3]) / np.asarray(vsize_xyz, dtype=np.float64)).astype(np.int32)

    def point_to_voxel(self, pc):
        pts = pc.numpy() if hasattr(pc, "numpy") else np.asarray(pc)
        pts = np.ascontiguousarray(pts, dtype=np.float32)
        n, stride = pts.shape
        nf, mv, mp = self.num_point_features, self.max_num_voxels, self.max_num_points_per_voxel
        voxels = np.empty((mv, mp, nf), dtype=np.float32)
        coords = np.empty((mv, 3), dtype=np.int32)
        num = np.empty((mv,), dtype=np.int32)
        count = ctypes.c_int(0)
        P = lambda a: a.ctypes.data_as(ctypes.c_void_p)
        _lib.call("crb3d_point_to_voxel_cpu", P(pts), n, stride, nf, P(self.coors_range), P(self.vsize),
                  P(self.grid_size), mp, mv, P(voxels), P(coords), P(num), ctypes.byref(count))
        m = count.value
        return tv.from_numpy(voxels[:m]), tv.from_numpy(coords[:m]), tv.from_numpy(num[:m])

    def point_to_voxel_empty_mean(self, pc):
        raise NotImplementedError("point_to_voxel_empty_mean is not used by the reference")
