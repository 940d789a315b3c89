"""Host-side mirror of the reference's box-op wrappers on the crb3d kernels: same names and argument meaning as
pcdet/ops/iou3d_nms/iou3d_nms_utils.py:12-116, pcdet/ops/roiaware_pool3d/roiaware_pool3d_utils.py:9-107 and
pcdet/models/model_utils/model_nms_utils.py:6-66. Unlike the reference, NMS results never leave the device."""
import torch
import torch.nn as nn
from torch.autograd import Function

from . import ops


# ---------------------------------------------------------------------------------------------- iou3d_nms_utils
def boxes_bev_iou_cpu(boxes_a, boxes_b):
    a = torch.as_tensor(boxes_a).float()
    b = torch.as_tensor(boxes_b).float()
    assert not (a.is_cuda or b.is_cuda), "Only support CPU tensors"
    out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32)
    ops.boxes_iou_bev_cpu(a[:, :7].contiguous(), b[:, :7].contiguous(), out)
    return out.numpy() if not isinstance(boxes_a, torch.Tensor) else out


def boxes_iou_bev(boxes_a, boxes_b):
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    return ops.boxes_iou_bev(boxes_a, boxes_b)


def boxes_iou3d_gpu(boxes_a, boxes_b):
    """(N,7) x (M,7) -> (N,M) 3-D IoU: BEV overlap (kernel) x height overlap / union volume (iou3d_nms_utils.py:48-81)."""
    assert boxes_a.shape[1] == boxes_b.shape[1] == 7
    a_max = (boxes_a[:, 2] + boxes_a[:, 5] / 2).view(-1, 1)
    a_min = (boxes_a[:, 2] - boxes_a[:, 5] / 2).view(-1, 1)
    b_max = (boxes_b[:, 2] + boxes_b[:, 5] / 2).view(1, -1)
    b_min = (boxes_b[:, 2] - boxes_b[:, 5] / 2).view(1, -1)
    overlaps_bev = ops.boxes_overlap_bev(boxes_a, boxes_b)
    overlaps_h = torch.clamp(torch.min(a_max, b_max) - torch.max(a_min, b_min), min=0)
    overlaps_3d = overlaps_bev * overlaps_h
    vol_a = (boxes_a[:, 3] * boxes_a[:, 4] * boxes_a[:, 5]).view(-1, 1)
    vol_b = (boxes_b[:, 3] * boxes_b[:, 4] * boxes_b[:, 5]).view(1, -1)
    return overlaps_3d / torch.clamp(vol_a + vol_b - overlaps_3d, min=1e-6)


def _nms(boxes, scores, thresh, rotated, pre_maxsize=None):
    assert boxes.shape[1] == 7
    order = scores.sort(0, descending=True)[1]
    if pre_maxsize is not None:
        order = order[:pre_maxsize]
    keep, num = ops.nms_sorted(boxes[order].contiguous(), thresh, rotated=rotated)
    return order[keep[: int(num.item())]].contiguous(), None


def nms_gpu(boxes, scores, thresh, pre_maxsize=None, **kwargs):
    return _nms(boxes, scores, thresh, True, pre_maxsize)


def nms_normal_gpu(boxes, scores, thresh, **kwargs):
    return _nms(boxes, scores, thresh, False)


def class_agnostic_nms(box_scores, box_preds, nms_config, score_thresh=None):
    """model_nms_utils.py:6-25. nms_config: object/dict with NMS_TYPE, NMS_THRESH, NMS_PRE_MAXSIZE, NMS_POST_MAXSIZE."""
    get = (lambda k: nms_config[k]) if isinstance(nms_config, dict) else (lambda k: getattr(nms_config, k))
    src_box_scores = box_scores
    if score_thresh is not None:
        scores_mask = box_scores >= score_thresh
        box_scores, box_preds = box_scores[scores_mask], box_preds[scores_mask]
    selected = []
    if box_scores.shape[0] > 0:
        box_scores_nms, indices = torch.topk(box_scores, k=min(get("NMS_PRE_MAXSIZE"), box_scores.shape[0]))
        fn = {"nms_gpu": nms_gpu, "nms_normal_gpu": nms_normal_gpu}[get("NMS_TYPE")]
        keep_idx, _ = fn(box_preds[indices][:, 0:7], box_scores_nms, get("NMS_THRESH"))
        selected = indices[keep_idx[: get("NMS_POST_MAXSIZE")]]
    if score_thresh is not None:
        selected = scores_mask.nonzero().view(-1)[selected]
    return selected, src_box_scores[selected]


# ---------------------------------------------------------------------------------------------- roiaware_pool3d_utils
def points_in_boxes_cpu(points, boxes):
    p = torch.as_tensor(points).float()
    b = torch.as_tensor(boxes).float()
    assert b.shape[1] == 7 and p.shape[1] == 3
    out = torch.zeros((b.shape[0], p.shape[0]), dtype=torch.int32)
    ops.points_in_boxes_cpu(b.contiguous(), p.contiguous(), out)
    return out.numpy() if not isinstance(points, torch.Tensor) else out


def points_in_boxes_gpu(points, boxes):
    """points (B, M, 3), boxes (B, T, 7) -> (B, M) int32 index of the first containing box, -1 = background."""
    assert boxes.shape[0] == points.shape[0] and boxes.shape[2] == 7 and points.shape[2] == 3
    return ops.points_in_boxes(boxes, points)


class RoIAwarePool3dFunction(Function):
    @staticmethod
    def forward(ctx, rois, pts, pts_feature, out_size, max_pts_each_voxel, pool_method):
        assert rois.shape[1] == 7 and pts.shape[1] == 3
        ox, oy, oz = (out_size,) * 3 if isinstance(out_size, int) else out_size
        n, C, P = rois.shape[0], pts_feature.shape[-1], pts.shape[0]
        pooled = pts_feature.new_zeros((n, ox, oy, oz, C))
        argmax = pts_feature.new_zeros((n, ox, oy, oz, C), dtype=torch.int)
        pts_idx_of_voxels = pts_feature.new_zeros((n, ox, oy, oz, max_pts_each_voxel), dtype=torch.int)
        method = {"max": 0, "avg": 1}[pool_method]
        ops.roiaware_pool3d_forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled, method)
        ctx.roiaware_pool3d_for_backward = (pts_idx_of_voxels, argmax, method, P, C)
        return pooled

    @staticmethod
    def backward(ctx, grad_out):
        pts_idx_of_voxels, argmax, method, P, C = ctx.roiaware_pool3d_for_backward
        grad_in = grad_out.new_zeros((P, C))
        ops.roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out.contiguous(), grad_in, method)
        return None, None, grad_in, None, None, None


class RoIAwarePool3d(nn.Module):
    def __init__(self, out_size, max_pts_each_voxel=128):
        super().__init__()
        self.out_size, self.max_pts_each_voxel = out_size, max_pts_each_voxel

    def forward(self, rois, pts, pts_feature, pool_method="max"):
        assert pool_method in ("max", "avg")
        return RoIAwarePool3dFunction.apply(rois, pts, pts_feature, self.out_size, self.max_pts_each_voxel, pool_method)
