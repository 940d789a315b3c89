"""Stand-in for `pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda` (pcdet/ops/roiaware_pool3d/src/roiaware_pool3d.cpp:172-177)."""
from crb3d import ops


def forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled_features, pool_method):
    ops.roiaware_pool3d_forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled_features, pool_method)
    return 1


def backward(pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method):
    ops.roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method)
    return 1


def points_in_boxes_gpu(boxes, pts, box_idx_of_points):
    ops.points_in_boxes(boxes, pts, out=box_idx_of_points)
    return 1


def points_in_boxes_cpu(boxes, pts, pts_indices):
    ops.points_in_boxes_cpu(boxes, pts, pts_indices)
    return 1
