"""Stand-in for `pcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda`
(pcdet/ops/pointnet2/pointnet2_stack/src/pointnet2_api.cpp:13-30). The PV-RCNN++ entry points (vector_pool*, the stacked
local-neighbour queries behind them) are outside the CRB hot path and raise NotImplementedError."""
from crb3d import ops


def ball_query_wrapper(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx):
    ops.ball_query(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx)
    return 1


def group_points_wrapper(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    ops.group_points(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out)
    return 1


def group_points_grad_wrapper(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    ops.group_points_grad(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features)
    return 1


def farthest_point_sampling_wrapper(b, n, m, points, temp, idx):
    ops.farthest_point_sampling(b, n, m, points, temp, idx)
    return 1


def stack_farthest_point_sampling_wrapper(points, temp, xyz_batch_cnt, idx, num_sampled_points):
    ops.stack_farthest_point_sampling(points, temp, xyz_batch_cnt, idx, num_sampled_points)
    return 1


def three_nn_wrapper(unknown, unknown_batch_cnt, known, known_batch_cnt, dist2, idx):
    ops.three_nn(unknown_batch_cnt.shape[0], unknown.shape[0], known.shape[0], unknown, unknown_batch_cnt, known,
                 known_batch_cnt, dist2, idx)


def three_interpolate_wrapper(features, idx, weight, out):
    ops.three_interpolate(out.shape[0], features.shape[1], features, idx, weight, out)


def three_interpolate_grad_wrapper(grad_out, idx, weight, grad_features):
    ops.three_interpolate_grad(grad_out.shape[0], grad_out.shape[1], grad_out, idx, weight, grad_features)


def voxel_query_wrapper(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords, point_indices, idx):
    ops.voxel_query(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords, point_indices, idx)
    return 1


def _not_on_path(name):
    def fn(*args, **kwargs):
        raise NotImplementedError("%s belongs to PV-RCNN++ and is outside the CRB hot path" % name)
    fn.__name__ = name
    return fn


query_stacked_local_neighbor_idxs_wrapper_stack = _not_on_path("query_stacked_local_neighbor_idxs_wrapper_stack")
query_three_nn_by_stacked_local_idxs_wrapper_stack = _not_on_path("query_three_nn_by_stacked_local_idxs_wrapper_stack")
vector_pool_wrapper = _not_on_path("vector_pool_wrapper")
vector_pool_grad_wrapper = _not_on_path("vector_pool_grad_wrapper")
