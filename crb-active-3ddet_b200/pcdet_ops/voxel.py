"""`pcdet.ops.voxel` - the device voxelizer named by the north star (no such package exists in the reference at this
commit): hard voxelization + MeanVFE on the GPU with the exact Point2VoxelCPU3d semantics."""
import torch

from crb3d import ops


def hard_voxelize(points, voxel_size, point_cloud_range, max_points_per_voxel, max_voxels_per_frame, batch_size=None):
    """points: CUDA (N, 1+C) with the batch index in column 0 (the collate layout of `batch_dict['points']`), sorted by
    batch. Returns (voxel_features (M,C) [MeanVFE], voxel_coords (M,4) int32 [b,z,y,x], voxel_num_points (M,))."""
    b = points[:, 0].long()
    if batch_size is None:
        batch_size = int(b.max().item()) + 1 if points.shape[0] else 1
    counts = torch.bincount(b, minlength=batch_size)
    offs = torch.zeros(batch_size + 1, dtype=torch.int32, device=points.device)
    offs[1:] = torch.cumsum(counts, 0)
    res = ops.voxelize(points, offs, batch_size, point_cloud_range, voxel_size, max_points_per_voxel,
                       max_voxels_per_frame, xyz_col=1, feat_col=1, n_feat=points.shape[1] - 1)
    return res["mean"], res["coords"], res["num_points"]
