"""Registers the crb3d stand-ins under the import names of the reference's compiled extension modules, so the unmodified
reference wrappers (`from . import iou3d_nms_cuda` in pcdet/ops/iou3d_nms/iou3d_nms_utils.py:8, `roiaware_pool3d_cuda`
in roiaware_pool3d_utils.py:6, `pointnet2_stack_cuda` in pointnet2_stack/pointnet2_utils.py:5, `pointnet2_batch_cuda` in
pointnet2_batch/pointnet2_utils.py:7, `roipoint_pool3d_cuda` in roipoint_pool3d_utils.py:6) and `import spconv` /
`cumm` resolve to this library. Call before importing pcdet:

    import sys; sys.path.insert(0, "<repo>/crb-active-3ddet_b200")
    import crb3d.dropin; crb3d.dropin.install()
"""
import importlib
import importlib.abc
import importlib.machinery
import sys

ALIASES = {
    "pcdet.ops.iou3d_nms.iou3d_nms_cuda": "pcdet_ops.iou3d_nms_cuda",
    "pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda": "pcdet_ops.roiaware_pool3d_cuda",
    "pcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda": "pcdet_ops.pointnet2_stack_cuda",
    "pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda": "pcdet_ops.pointnet2_batch_cuda",
    "pcdet.ops.roipoint_pool3d.roipoint_pool3d_cuda": "pcdet_ops.roipoint_pool3d_cuda",
    "pcdet.ops.voxel": "pcdet_ops.voxel",
    # build_strategy('crb', ...) of pcdet/query_strategies/__init__.py:13-29 then returns the crb3d CRBSampling
    "pcdet.query_strategies.crb_sampling": "crb3d.crb_strategy",
}
EXPECTED = {
    "pcdet.ops.iou3d_nms.iou3d_nms_cuda": ["boxes_overlap_bev_gpu", "boxes_iou_bev_gpu", "nms_gpu", "nms_normal_gpu",
                                           "boxes_iou_bev_cpu"],
    "pcdet.ops.roiaware_pool3d.roiaware_pool3d_cuda": ["forward", "backward", "points_in_boxes_gpu", "points_in_boxes_cpu"],
    "pcdet.ops.pointnet2.pointnet2_stack.pointnet2_stack_cuda": [
        "ball_query_wrapper", "voxel_query_wrapper", "farthest_point_sampling_wrapper",
        "stack_farthest_point_sampling_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "three_nn_wrapper",
        "three_interpolate_wrapper", "three_interpolate_grad_wrapper", "query_stacked_local_neighbor_idxs_wrapper_stack",
        "query_three_nn_by_stacked_local_idxs_wrapper_stack", "vector_pool_wrapper", "vector_pool_grad_wrapper"],
    "pcdet.ops.pointnet2.pointnet2_batch.pointnet2_batch_cuda": [
        "ball_query_wrapper", "group_points_wrapper", "group_points_grad_wrapper", "gather_points_wrapper",
        "gather_points_grad_wrapper", "farthest_point_sampling_wrapper", "three_nn_wrapper", "three_interpolate_wrapper",
        "three_interpolate_grad_wrapper"],
    "pcdet.ops.roipoint_pool3d.roipoint_pool3d_cuda": ["forward"],
    "pcdet.ops.voxel": ["hard_voxelize"],
    "pcdet.query_strategies.crb_sampling": ["CRBSampling"],
}


class _AliasFinder(importlib.abc.MetaPathFinder, importlib.abc.Loader):
    """`from . import iou3d_nms_cuda` inside the reference package ends in an import of the fully qualified name; this
    finder answers it with the stand-in module."""

    def find_spec(self, fullname, path, target=None):
        if fullname in ALIASES:
            return importlib.machinery.ModuleSpec(fullname, self)
        return None

    def create_module(self, spec):
        return importlib.import_module(ALIASES[spec.name])

    def exec_module(self, module):
        pass


_finder = None


def install():
    """Idempotent. Returns the list of aliased module names."""
    global _finder
    if _finder is None:
        _finder = _AliasFinder()
        sys.meta_path.insert(0, _finder)
    for alias, real in ALIASES.items():
        sys.modules[alias] = importlib.import_module(real)
    return sorted(ALIASES)


def accelerate_bev_backbone(module):
    """Gives an instance of the reference's `BaseBEVBackbone` (pcdet/models/backbones_2d/base_bev_backbone.py:7-112; same
    `blocks` / `deblocks` structure and parameter names as crb3d.second.BaseBEVBackbone) the tensor-core inference plan of
    this library: in eval mode without autograd `forward(data_dict)` runs the folded-BatchNorm plan (3x3 convs on
    crb3d_bev_conv3x3_tf32, deblocks on crb3d_bev_gemm_tf32 writing their slice of the concatenated map); training and
    autograd keep the module's own forward. Call again (or `module.build_inference_plan()`) after loading a checkpoint."""
    import types

    import torch

    from .second import BaseBEVBackbone as Ours
    original_forward = module.forward
    module._fold = Ours._fold
    module._tc_conv_pays = Ours._tc_conv_pays
    module._build_sparse_fills = Ours._build_sparse_fills
    module.build_inference_plan = types.MethodType(Ours.build_inference_plan, module)
    module.forward_inference = types.MethodType(Ours.forward_inference, module)

    def forward(self, data_dict):
        x = data_dict["spatial_features"]
        if not self.training and not torch.is_grad_enabled() and getattr(self, "_plan", None) is not None and x.is_cuda:
            # `_occupancy` = (coords, n_dev) of the sparse tensor behind spatial_features, when the caller provides it: block 1 then
            # runs in sparse-tile mode (crb3d.second.BaseBEVBackbone.forward_inference)
            data_dict["spatial_features_2d"] = self.forward_inference(x, None, data_dict.get("_occupancy"))
            return data_dict
        return original_forward(data_dict)

    module.forward = types.MethodType(forward, module)
    module.build_inference_plan()
    return module
