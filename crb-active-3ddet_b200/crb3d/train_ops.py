"""Anchor-head training path on the device: target assignment and the anchor-head losses with their
gradients, mirroring the reference's interfaces:

  AxisAlignedTargetAssigner  pcdet/models/dense_heads/target_assigner/axis_aligned_target_assigner.py:8-210
  anchor_head_loss           AnchorHeadTemplate.get_loss, pcdet/models/dense_heads/anchor_head_template.py:101-229
                             (loss_utils.SigmoidFocalClassificationLoss / WeightedSmoothL1Loss / WeightedCrossEntropyLoss)

CUDA tensors only (no CPU path); kernels in csrc/train_ops.cu."""
import ctypes

import torch

from . import _lib, ops
from .ops import _f32c, _need_cuda, _p, _stream, _ws, _ws_bytes


class AxisAlignedTargetAssigner(object):
    """Same role and return value as the reference class; built from the anchor configuration instead of the yaml node.

    anchor_cfgs: the ANCHOR_GENERATOR_CONFIG list (class_name, anchor_sizes, anchor_rotations, anchor_bottom_heights,
    matched_threshold, unmatched_threshold per class); class_names: the detector's CLASS_NAMES (label = index + 1)."""

    def __init__(self, anchor_cfgs, class_names, match_height=False, pos_fraction=-1.0):
        if match_height:
            raise NotImplementedError("match_height=True (3-D IoU matching) is not used by the reference configs")
        if pos_fraction is not None and pos_fraction >= 0:
            raise NotImplementedError("POS_FRACTION >= 0 (random anchor sampling) is not used by the reference configs")
        tc, mt, ut = [], [], []
        for cfg in anchor_cfgs:
            n = len(cfg["anchor_sizes"]) * len(cfg["anchor_rotations"]) * len(cfg["anchor_bottom_heights"])
            tc += [list(class_names).index(cfg["class_name"]) + 1] * n
            mt += [float(cfg["matched_threshold"])] * n
            ut += [float(cfg["unmatched_threshold"])] * n
        self.n_types = len(tc)
        self._tc = (ctypes.c_int * self.n_types)(*tc)
        self._mt = (ctypes.c_float * self.n_types)(*mt)
        self._ut = (ctypes.c_float * self.n_types)(*ut)

    def assign_targets(self, anchors, gt_boxes_with_classes):
        """anchors (A,7) float32 CUDA in the head's (y, x, type) order; gt_boxes_with_classes (B, M, 8).
        Returns the reference's dict: box_cls_labels (B,A) int32, box_reg_targets (B,A,7), reg_weights (B,A) - plus num_pos (B)."""
        _need_cuda(anchors, gt_boxes_with_classes)
        anchors, gt = _f32c(anchors.reshape(-1, anchors.shape[-1])[:, :7]), _f32c(gt_boxes_with_classes)
        A, (B, M, S) = anchors.shape[0], gt.shape
        dev = anchors.device
        labels = torch.empty((B, A), dtype=torch.int32, device=dev)
        reg_t = torch.empty((B, A, 7), dtype=torch.float32, device=dev)
        reg_w = torch.empty((B, A), dtype=torch.float32, device=dev)
        num_pos = torch.empty((B,), dtype=torch.int32, device=dev)
        ws = _ws(_ws_bytes("crb3d_assign_targets_workspace_bytes", B, A, M), dev)
        _lib.call("crb3d_assign_targets_axis_aligned", _p(anchors), A, self.n_types, self._tc, self._mt, self._ut, _p(gt), B, M, S,
                  _p(labels), _p(reg_t), _p(reg_w), _p(num_pos), _p(ws), ws.numel(), _stream(dev))
        return {"box_cls_labels": labels, "box_reg_targets": reg_t, "reg_weights": reg_w, "num_pos": num_pos}


class _AnchorHeadLoss(torch.autograd.Function):
    @staticmethod
    def forward(ctx, cls_preds, box_preds, dir_preds, labels, reg_targets, anchors, cfg):
        B, A, n_class = cls_preds.shape
        dev = cls_preds.device
        cls_preds, box_preds = _f32c(cls_preds), _f32c(box_preds)
        dir_preds = _f32c(dir_preds) if dir_preds is not None else None
        nb = dir_preds.shape[-1] if dir_preds is not None else 0
        losses = torch.empty((3,), dtype=torch.float32, device=dev)
        g_cls, g_box = torch.empty_like(cls_preds), torch.empty_like(box_preds)
        g_dir = torch.empty_like(dir_preds) if dir_preds is not None else None
        cw = cfg.get("code_weights")
        cw = (ctypes.c_float * 7)(*[float(x) for x in cw[:7]]) if cw is not None else None
        lw = (ctypes.c_float * 3)(float(cfg["cls_weight"]), float(cfg["loc_weight"]), float(cfg["dir_weight"]))
        ws = _ws(_ws_bytes("crb3d_anchor_head_loss_workspace_bytes", B, A), dev)
        _lib.call("crb3d_anchor_head_loss", _p(cls_preds), _p(box_preds), _p(dir_preds), _p(labels), _p(reg_targets), _p(anchors), B, A,
                  n_class, nb, cw, float(cfg.get("alpha", 0.25)), float(cfg.get("gamma", 2.0)), float(cfg.get("beta", 1.0 / 9.0)),
                  float(cfg.get("dir_offset", 0.78539)), lw, _p(losses), _p(g_cls), _p(g_box), _p(g_dir), _p(ws), ws.numel(),
                  _stream(dev))
        ctx.save_for_backward(g_cls, g_box, g_dir if g_dir is not None else torch.empty(0, device=dev))
        ctx.has_dir = g_dir is not None
        return losses

    @staticmethod
    def backward(ctx, grad_losses):
        g_cls, g_box, g_dir = ctx.saved_tensors
        # the three losses enter the total with the same upstream factor in every caller (rpn_loss = cls + loc + dir); general
        # upstream gradients are honoured per component
        gc, gl, gd = grad_losses[0], grad_losses[1], grad_losses[2]
        return (g_cls * gc, g_box * gl, (g_dir * gd) if ctx.has_dir else None, None, None, None, None)


def anchor_head_loss(cls_preds, box_preds, dir_preds, labels, reg_targets, anchors, cfg):
    """(rpn_loss_cls, rpn_loss_loc, rpn_loss_dir) as a (3,) tensor with autograd to the three prediction tensors; their sum is the
    reference's rpn_loss. cfg: dict(cls_weight, loc_weight, dir_weight[, code_weights, alpha, gamma, beta, dir_offset])
    (LOSS_CONFIG.LOSS_WEIGHTS + DIR_OFFSET). Shapes: cls (B,A,n_class), box (B,A,7), dir (B,A,bins) | None, labels (B,A) int32,
    reg_targets (B,A,7), anchors (A,7)."""
    _need_cuda(cls_preds, box_preds, labels, reg_targets, anchors)
    ops._check(labels=(labels, torch.int32), reg_targets=(reg_targets, torch.float32), anchors=(anchors, torch.float32))
    return _AnchorHeadLoss.apply(cls_preds, box_preds, dir_preds, labels, reg_targets, anchors, cfg)
