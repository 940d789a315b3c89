"""Stand-in for `pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda`
(pcdet/ops/pointnet2/pointnet2_batch/src/pointnet2_api.cpp:10-24), the (B, N, 3) / (B, C, N) layout used by PointRCNN's
PointNet2MSG backbone. Same names and argument order as the pybind module; CUDA tensors only."""
from crb3d import ops


def ball_query_wrapper(b, n, m, radius, nsample, new_xyz, xyz, idx):
    ops.ball_query_batch(b, n, m, radius, nsample, new_xyz, xyz, idx)
    return 1


def group_points_wrapper(b, c, n, npoints, nsample, points, idx, out):
    ops.group_points_batch(b, c, n, npoints, nsample, points, idx, out)
    return 1


def group_points_grad_wrapper(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    ops.group_points_grad_batch(b, c, n, npoints, nsample, grad_out, idx, grad_points)
    return 1


def gather_points_wrapper(b, c, n, npoints, points, idx, out):
    # sampling_gpu.cu:15-33: out (b, c, npoints) = points[b, c, idx[b, :]] - the nsample = 1 case of group_points
    ops.group_points_batch(b, c, n, npoints, 1, points, idx, out)
    return 1


def gather_points_grad_wrapper(b, c, n, npoints, grad_out, idx, grad_points):
    ops.group_points_grad_batch(b, c, n, npoints, 1, grad_out, idx, grad_points)
    return 1


def farthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    ops.farthest_point_sampling(b, n, m, points, temp, idx)
    return 1


def three_nn_wrapper(b, n, m, unknown, known, dist2, idx):
    ops.three_nn_batch(b, n, m, unknown, known, dist2, idx)


def three_interpolate_wrapper(b, c, m, n, points, idx, weight, out):
    ops.three_interpolate_batch(b, c, m, n, points, idx, weight, out)


def three_interpolate_grad_wrapper(b, c, n, m, grad_out, idx, weight, grad_points):
    ops.three_interpolate_grad_batch(b, c, n, m, grad_out, idx, weight, grad_points)
