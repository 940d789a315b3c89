"""Stand-in for `pcdet.ops.roipoint_pool3d.roipoint_pool3d_cuda` (pcdet/ops/roipoint_pool3d/src/roipoint_pool3d.cpp:23-51),
the canonical-box point pooling of PointRCNN's RoI head. CUDA tensors only."""
from crb3d import ops


def forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag):
    ops.roipoint_pool3d_forward(xyz, boxes3d, pts_feature, pooled_features, pooled_empty_flag)
    return 1
