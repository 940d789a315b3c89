"""Stand-in for the pybind module `pcdet.ops.iou3d_nms.iou3d_nms_cuda` (pcdet/ops/iou3d_nms/src/iou3d_nms_api.cpp:12-16).
Outputs are caller-allocated and written in place; `keep` is a CPU LongTensor as in iou3d_nms.cpp:90-136."""
import torch

from crb3d import ops


def _check(t, name):
    if not t.is_cuda:
        raise RuntimeError("%s must be a CUDA tensor" % name)
    if not t.is_contiguous():
        raise RuntimeError("%s must be contiguous" % name)


def boxes_overlap_bev_gpu(boxes_a, boxes_b, ans_overlap):
    _check(boxes_a, "boxes_a"); _check(boxes_b, "boxes_b"); _check(ans_overlap, "ans_overlap")
    ops.boxes_overlap_bev(boxes_a, boxes_b, out=ans_overlap)
    return 1


def boxes_iou_bev_gpu(boxes_a, boxes_b, ans_iou):
    _check(boxes_a, "boxes_a"); _check(boxes_b, "boxes_b"); _check(ans_iou, "ans_iou")
    ops.boxes_iou_bev(boxes_a, boxes_b, out=ans_iou)
    return 1


def _nms(boxes, keep, thresh, rotated):
    _check(boxes, "boxes")
    if keep.is_cuda or keep.dtype != torch.int64:
        raise RuntimeError("keep must be a CPU LongTensor")
    k, n = ops.nms_sorted(boxes, thresh, rotated=rotated)
    num = int(n.item())
    keep[:num] = k[:num].cpu()
    return num


def nms_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, True)


def nms_normal_gpu(boxes, keep, nms_overlap_thresh):
    return _nms(boxes, keep, nms_overlap_thresh, False)


def boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou):
    ops.boxes_iou_bev_cpu(boxes_a, boxes_b, ans_iou)
    return 1
