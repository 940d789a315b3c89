"""ctypes loader for libcrb3d_sm100.so (the C ABI declared in include/crb3d.h).

There is no CPU fallback: if the library is missing or a call returns an error code, a RuntimeError is raised.
"""
import ctypes
import os
from ctypes import c_char_p, c_double, c_float, c_int, c_int64, c_size_t, c_void_p, POINTER

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(os.path.dirname(_HERE), "lib", "libcrb3d_sm100.so")

P = c_void_p  # every device/host buffer crosses as a raw address

# name -> argtypes (restype is int unless listed in _RESTYPES)
SIGNATURES = {
    "crb3d_version": [],
    "crb3d_strerror": [c_int],
    "crb3d_diag_init": [],
    "crb3d_last_device_error": [P],
    "crb3d_diag_clear": [],
    "crb3d_device_sm_count": [P],
    "crb3d_voxelize_workspace_bytes": [c_int64, c_int, c_int, POINTER(c_size_t)],
    "crb3d_voxelize": [P, c_int64, c_int, c_int, c_int, c_int, P, c_int, P, P, P, c_int, c_int, P, P, P, P, P, P,
                       c_size_t, P],
    "crb3d_point_to_voxel_cpu": [P, c_int64, c_int, c_int, P, P, P, c_int, c_int, P, P, P, P],
    "crb3d_subm_rulebook_workspace_bytes": [c_int, POINTER(c_size_t)],
    "crb3d_subm_rulebook": [P, c_int, P, P, P, P, P, P, c_size_t, P],
    "crb3d_conv_out_shape": [P, P, P, P, P, P],
    "crb3d_sparse_rulebook_workspace_bytes": [c_int, P, POINTER(c_size_t)],
    "crb3d_sparse_rulebook_coords": [P, c_int, P, c_int, P, P, P, P, P, P, P, c_int, P, P, c_size_t, P],
    "crb3d_sparse_rulebook_pairs": [P, c_int, P, c_int, P, P, P, P, P, P, c_int, P, P, P, c_size_t, P],
    "crb3d_rulebook_compact_pairs_workspace_bytes": [c_int, c_int, POINTER(c_size_t)],
    "crb3d_rulebook_compact_pairs": [P, c_int, c_int, c_int, P, P, P, P, c_size_t, P],
    "crb3d_spconv_forward_f32": [P, P, P, c_int, c_int, c_int, c_int, c_int64, c_int64, c_int64, P, P, P, c_int, P, P, P],
    "crb3d_spconv_forward_tf32": [P, c_int, P, P, c_int, c_int, c_int, c_int, P, P, P, c_int, P, P, P],
    "crb3d_spconv_wgrad_workspace_bytes": [c_int, c_int, c_int, c_int, POINTER(c_size_t)],
    "crb3d_spconv_wgrad_f32": [P, P, P, c_int, c_int, c_int, c_int, c_int, P, P, c_size_t, P],
    "crb3d_sparse_to_dense": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P],
    "crb3d_dense_to_sparse": [P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P],
    "crb3d_boxes_overlap_bev": [P, c_int, P, c_int, P, P],
    "crb3d_boxes_iou_bev": [P, c_int, P, c_int, P, P],
    "crb3d_nms_workspace_bytes": [c_int, POINTER(c_size_t)],
    "crb3d_nms": [P, c_int, c_float, c_int, c_int, P, P, P, c_size_t, P],
    "crb3d_nms_batched_workspace_bytes": [c_int, c_int, POINTER(c_size_t)],
    "crb3d_nms_batched": [P, P, c_int, c_int, c_float, c_int, c_int, P, c_int, P, P, c_size_t, P],
    "crb3d_nms_mask": [P, c_int, c_float, c_int, P, P, c_size_t, P],
    "crb3d_boxes_iou_bev_cpu": [P, c_int, P, c_int, P],
    "crb3d_points_in_boxes": [P, P, c_int, c_int, c_int, P, P],
    "crb3d_points_in_boxes_stack": [P, c_int, P, c_int, P, P, c_int, c_int, c_int, P, P, P, P],
    "crb3d_points_in_boxes_ranges": [P, c_int, P, P, c_int, P, P, P, c_int, c_int, P, P, P, P],
    "crb3d_points_in_boxes_cpu": [P, c_int, P, c_int, P],
    "crb3d_roiaware_pool3d_forward": [P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P, P, P, c_int, P],
    "crb3d_roiaware_pool3d_backward": [P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, P],
    "crb3d_ball_query_stack": [c_int, c_int, c_float, c_int, P, P, P, P, P, c_int, P],
    "crb3d_group_points_stack": [c_int, c_int, c_int, c_int, P, P, P, P, P, P],
    "crb3d_group_points_grad_stack": [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P],
    "crb3d_farthest_point_sampling": [c_int, c_int, c_int, P, P, P, P],
    "crb3d_stack_farthest_point_sampling": [c_int, c_int, P, P, P, P, P, P],
    "crb3d_three_nn_stack": [c_int, c_int, c_int, P, P, P, P, P, P, P],
    "crb3d_three_interpolate_stack": [c_int, c_int, P, P, P, P, P],
    "crb3d_three_interpolate_grad_stack": [c_int, c_int, P, P, P, P, P],
    "crb3d_bev_gemm_tf32": [P, c_int64, c_int, c_int64, P, c_int, c_int, P, c_int, c_int, P, P, P, P, c_int, c_int, c_int, P],
    "crb3d_bev_conv3x3_tf32": [P, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, P],
    "crb3d_bev_conv_gemm_tf32": [P, c_int, c_int, c_int, c_int, P, c_int, c_int, c_int, c_int, P, c_int, P, P],
    "crb3d_mask_collate_points_workspace_bytes": [c_int64, POINTER(c_size_t)],
    "crb3d_mask_collate_points": [P, c_int64, c_int, c_int, P, c_int, P, P, P, P, c_size_t, P],
    "crb3d_furthest_first_workspace_bytes": [c_int, POINTER(c_size_t)],
    "crb3d_furthest_first": [P, c_int, c_int, P, c_int, P, P, c_size_t, P],
    "crb3d_subm_rulebook_cellmap": [P, c_int, P, P, P, P, P, P, P, P],
    "crb3d_sparse_rulebook_cellmap": [c_int, P, POINTER(c_size_t), POINTER(c_size_t), POINTER(c_int64)],
    "crb3d_assign_targets_workspace_bytes": [c_int, c_int64, c_int, POINTER(c_size_t)],
    "crb3d_assign_targets_axis_aligned": [P, c_int64, c_int, P, P, P, P, c_int, c_int, c_int, P, P, P, P, P, c_size_t, P],
    "crb3d_anchor_head_loss_workspace_bytes": [c_int, c_int64, POINTER(c_size_t)],
    "crb3d_anchor_head_loss": [P, P, P, P, P, P, c_int, c_int64, c_int, c_int, P, c_float, c_float, c_float, c_float, P, P, P, P, P, P,
                               c_size_t, P],
    "crb3d_query_stacked_local_neighbor_idxs_workspace_bytes": [c_int, POINTER(c_size_t)],
    "crb3d_query_stacked_local_neighbor_idxs": [P, P, P, P, c_int, c_int, P, P, P, c_int, c_float, c_int, c_int, P, c_size_t, P],
    "crb3d_query_three_nn_by_stacked_local_idxs": [P, P, P, P, P, P, c_int, c_int, P],
    "crb3d_vector_pool_stack": [P, P, P, P, P, c_int, c_int, c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, c_int, c_int,
                                P, P, P, P, P, P],
    "crb3d_vector_pool_grad_stack": [P, P, P, P, c_int, c_int, c_int, c_int, P],
    "crb3d_voxel_query_stack": [c_int, c_int, c_int, c_int, c_int, c_float, c_int, c_int, c_int, P, P, P, P, P, P],
    "crb3d_ball_query_batch": [c_int, c_int, c_int, c_float, c_int, P, P, P, P],
    "crb3d_group_points_batch": [c_int, c_int, c_int, c_int, c_int, P, P, P, P],
    "crb3d_group_points_grad_batch": [c_int, c_int, c_int, c_int, c_int, P, P, P, P],
    "crb3d_three_nn_batch": [c_int, c_int, c_int, P, P, P, P, P],
    "crb3d_three_interpolate_batch": [c_int, c_int, c_int, c_int, P, P, P, P, P],
    "crb3d_three_interpolate_grad_batch": [c_int, c_int, c_int, c_int, P, P, P, P, P],
    "crb3d_roipoint_pool3d_workspace_bytes": [c_int, c_int, c_int, POINTER(c_size_t)],
    "crb3d_roipoint_pool3d_forward": [c_int, c_int, c_int, c_int, c_int, P, P, P, P, P, P, c_size_t, P],
    "crb3d_sa_group_mlp_maxpool": [c_int, P, P, P, c_int, P, P, c_int, P, c_int, c_int, P, P, P, c_int, P],
    "crb3d_fc_gemm_workspace_bytes": [c_int64, c_int, c_int, POINTER(c_size_t)],
    "crb3d_fc_gemm_tf32": [P, c_int64, c_int, c_int64, P, c_int, P, P, c_int, P, P, c_size_t, P],
    "crb3d_bev_conv3x3_num_tiles": [c_int, c_int, c_int, POINTER(c_int)],
    "crb3d_bev_tile_plan_workspace_bytes": [c_int, c_int, c_int, POINTER(c_size_t)],
    "crb3d_bev_tile_plan": [P, c_int, P, c_int, c_int, c_int, c_int, P, P, P, P, P, c_size_t, P],
    "crb3d_bev_conv3x3_tf32_tiles": [P, c_int, c_int, c_int, c_int, P, c_int, P, c_int, P, P, P, P, P, P],
    "crb3d_bev_conv3x3_trace": [P, c_int],
    "crb3d_anchor_head_scores": [P, c_int64, c_int, P, P, P],
    "crb3d_anchor_head_scores_topk": [P, c_int, c_int64, c_int, c_float, c_int, P, P, P, P, P, P, P, P],
    "crb3d_anchor_decode_select": [P, P, P, c_int, c_int, c_int64, P, P, P],
    "crb3d_gather_rows_f32": [P, P, P, c_int, c_int, c_int64, c_int, c_float, P, P],
    "crb3d_gather_rows_i32": [P, P, P, c_int, c_int, c_int64, c_int, c_int, P, P],
    "crb3d_label_entropy": [P, P, c_int, c_int, P, P, P],
    "crb3d_label_entropy_ranges": [P, P, P, c_int, c_int, P, P, P],
    "crb3d_pairwise_sqdist_f64": [P, c_int, c_int, P, P],
    "crb3d_kde_greedy_workspace_bytes": [c_int, c_int, POINTER(c_size_t)],
    "crb3d_kde_greedy": [P, P, P, c_int, c_int, P, P, c_double, c_int, P, P, P, c_size_t, P],
}
_RESTYPES = {"crb3d_version": c_char_p, "crb3d_strerror": c_char_p}

_lib = None


def load():
    """Returns the loaded CDLL; raises RuntimeError (never falls back) when the library is absent."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            "libcrb3d_sm100.so not found at %s - build it with `python crb-active-3ddet_b200/build.py` "
            "(there is no CPU fallback for this path)" % LIB_PATH)
    lib = ctypes.CDLL(LIB_PATH)
    for name, argtypes in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here means header / library drift: fail loudly
        fn.argtypes = argtypes
        fn.restype = _RESTYPES.get(name, c_int)
    _lib = lib
    return lib


def check(code, what):
    if code != 0:
        msg = load().crb3d_strerror(code).decode()
        raise RuntimeError("%s failed: %s (code %d)" % (what, msg, code))


# CUDA kernels launched by one call of each entry point (read off csrc/*.cu; memsets are not counted). Used by
# bench.py to report `gpu_launches` = kernels of THIS library launched inside the timed region.
KERNELS_PER_CALL = {
    "crb3d_voxelize": 12, "crb3d_subm_rulebook": 2, "crb3d_sparse_rulebook_coords": 5, "crb3d_subm_rulebook_cellmap": 1, "crb3d_sparse_rulebook_pairs": 1,
    "crb3d_rulebook_compact_pairs": 6, "crb3d_spconv_forward_f32": 1, "crb3d_spconv_wgrad_f32": 2,
    "crb3d_sparse_to_dense": 1, "crb3d_dense_to_sparse": 1, "crb3d_boxes_overlap_bev": 1, "crb3d_boxes_iou_bev": 1,
    "crb3d_nms": 7, "crb3d_nms_batched": 7, "crb3d_nms_mask": 3, "crb3d_points_in_boxes": 1,
    "crb3d_points_in_boxes_stack": 2, "crb3d_points_in_boxes_ranges": 2, "crb3d_roiaware_pool3d_forward": 2,
    "crb3d_roiaware_pool3d_backward": 1, "crb3d_ball_query_stack": 1, "crb3d_group_points_stack": 1,
    "crb3d_group_points_grad_stack": 1, "crb3d_farthest_point_sampling": 1, "crb3d_stack_farthest_point_sampling": 1,
    "crb3d_three_nn_stack": 1, "crb3d_three_interpolate_stack": 1, "crb3d_three_interpolate_grad_stack": 1,
    "crb3d_label_entropy": 1, "crb3d_label_entropy_ranges": 1, "crb3d_pairwise_sqdist_f64": 1,
    "crb3d_anchor_head_scores": 1, "crb3d_anchor_head_scores_topk": 2, "crb3d_anchor_decode_select": 1, "crb3d_gather_rows_f32": 1, "crb3d_gather_rows_i32": 1,
    "crb3d_spconv_forward_tf32": 1, "crb3d_bev_gemm_tf32": 1, "crb3d_bev_conv3x3_tf32": 1, "crb3d_bev_conv_gemm_tf32": 1, "crb3d_sa_group_mlp_maxpool": 1, "crb3d_fc_gemm_tf32": 2, "crb3d_mask_collate_points": 6,
    "crb3d_voxel_query_stack": 1, "crb3d_ball_query_batch": 1, "crb3d_group_points_batch": 1, "crb3d_group_points_grad_batch": 1,
    "crb3d_three_nn_batch": 1, "crb3d_three_interpolate_batch": 1, "crb3d_three_interpolate_grad_batch": 1,
    "crb3d_roipoint_pool3d_forward": 1, "crb3d_assign_targets_axis_aligned": 2, "crb3d_anchor_head_loss": 3,
    "crb3d_bev_tile_plan": 5, "crb3d_bev_conv3x3_tf32_tiles": 2,
    "crb3d_query_stacked_local_neighbor_idxs": 8, "crb3d_query_three_nn_by_stacked_local_idxs": 1, "crb3d_vector_pool_stack": 1,
    "crb3d_vector_pool_grad_stack": 1,
}
LAUNCHES = {"kernels": 0, "calls": 0}


KERNEL_NAMES = {1: "spconv_fwd_tc", 2: "bev_conv3x3_tc", 3: "bev_conv3x3_pair_tc", 4: "bev_gemm_tc", 5: "hash_insert",
                6: "hash_find", 7: "bev_conv3x3_s2_tc", 8: "fc_gemm_tc"}
SITE_NAMES = {1: "full_a", 2: "empty_a", 3: "full_b", 4: "empty_b", 5: "acc_full", 6: "acc_empty", 7: "acc_bar", 8: "full_bar",
              9: "empty_bar", 10: "b_full"}
_DIAG_DEVICES = set()


def init_device(index):
    """Registers the diagnostics record with every kernel module on CUDA device `index` (idempotent; outside capture)."""
    if index in _DIAG_DEVICES:
        return
    import torch
    with torch.cuda.device(index):
        check(load().crb3d_diag_init(), "crb3d_diag_init")
    _DIAG_DEVICES.add(index)


def last_device_error():
    """None, or a dict describing the bounded wait / probe that gave up (readable even after the CUDA context died)."""
    out = (ctypes.c_uint * 12)()
    if load().crb3d_last_device_error(out) == 0:
        return None
    return dict(kernel=KERNEL_NAMES.get(out[1], out[1]), site=SITE_NAMES.get(out[2], out[2]), parity=out[3],
                block=(out[4], out[5]), thread=out[6], extra=out[7], waited_ms=((out[9] << 32) | out[8]) / 1e6, device=out[10])


def raise_if_device_error(prefix=""):
    e = last_device_error()
    if e is not None:
        raise RuntimeError("%scrb3d device error: %r" % (prefix, e))


def call(name, *args):
    """Calls a C-ABI entry point and raises on a non-zero status."""
    check(getattr(load(), name)(*args), name)
    LAUNCHES["calls"] += 1
    LAUNCHES["kernels"] += KERNELS_PER_CALL.get(name, 0)
