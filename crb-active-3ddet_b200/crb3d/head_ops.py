"""Torch front end for the anchor-head post-processing kernels (csrc/head.cu)."""
import ctypes

import numpy as np
import torch

from . import _lib
from .ops import _f32c, _need_cuda, _p, _stream

MAX_TYPES = 16


class AnchorSpec(ctypes.Structure):
    """Mirror of `struct AnchorSpec` in csrc/head.cu (native alignment)."""
    _fields_ = [
        ("nx", ctypes.c_int), ("ny", ctypes.c_int), ("n_types", ctypes.c_int),
        ("x0", ctypes.c_float), ("y0", ctypes.c_float),
        ("x_stride", ctypes.c_double), ("y_stride", ctypes.c_double),
        ("size", (ctypes.c_float * 3) * MAX_TYPES),
        ("rot", ctypes.c_float * MAX_TYPES),
        ("zc", ctypes.c_float * MAX_TYPES),
        ("dir_offset", ctypes.c_float), ("dir_limit_offset", ctypes.c_float),
        ("num_dir_bins", ctypes.c_int),
    ]


def make_anchor_spec(anchor_cfgs, pc_range, feature_map_size_xy, dir_offset=0.78539, dir_limit_offset=0.0, num_dir_bins=2):
    """anchor_cfgs: the ANCHOR_GENERATOR_CONFIG list (one entry per class, tools/cfgs/kitti_models/second.yaml:39-69).
    Reproduces AnchorGenerator.generate_anchors (anchor_generator.py:18-62) with align_center=False."""
    nx, ny = int(feature_map_size_xy[0]), int(feature_map_size_xy[1])
    s = AnchorSpec()
    s.nx, s.ny = nx, ny
    s.x0, s.y0 = float(pc_range[0]), float(pc_range[1])
    s.x_stride = (float(pc_range[3]) - float(pc_range[0])) / (nx - 1)
    s.y_stride = (float(pc_range[4]) - float(pc_range[1])) / (ny - 1)
    t = 0
    for cfg in anchor_cfgs:
        assert not cfg.get("align_center", False), "align_center anchors are not used by the reference configs"
        for size in cfg["anchor_sizes"]:
            for rot in cfg["anchor_rotations"]:
                for h in cfg["anchor_bottom_heights"]:
                    assert t < MAX_TYPES
                    for j in range(3):
                        s.size[t][j] = float(size[j])
                    s.rot[t] = float(rot)
                    s.zc[t] = float(np.float32(h) + np.float32(size[2]) / np.float32(2))
                    t += 1
    s.n_types = t
    s.dir_offset, s.dir_limit_offset, s.num_dir_bins = float(dir_offset), float(dir_limit_offset), int(num_dir_bins)
    return s


def anchors_tensor(spec):
    """All anchors (ny*nx*n_types, 7) float32 on the CPU, in the reference's order (for tests / the CPU oracle)."""
    xs = (np.float64(spec.x0) + spec.x_stride * np.arange(spec.nx)).astype(np.float32)
    ys = (np.float64(spec.y0) + spec.y_stride * np.arange(spec.ny)).astype(np.float32)
    out = np.zeros((spec.ny, spec.nx, spec.n_types, 7), np.float32)
    out[..., 0] = xs[None, :, None]
    out[..., 1] = ys[:, None, None]
    for t in range(spec.n_types):
        out[:, :, t, 2] = spec.zc[t]
        out[:, :, t, 3:6] = [spec.size[t][0], spec.size[t][1], spec.size[t][2]]
        out[:, :, t, 6] = spec.rot[t]
    return torch.from_numpy(out.reshape(-1, 7))


def anchor_head_scores(cls_preds, n_class):
    """cls_preds (..., n_class) channels-last logits -> (score = max sigmoid, label = argmax+1) flattened per anchor."""
    _need_cuda(cls_preds)
    cls_preds = _f32c(cls_preds)
    n = cls_preds.numel() // n_class
    score = torch.empty((n,), dtype=torch.float32, device=cls_preds.device)
    label = torch.empty((n,), dtype=torch.int32, device=cls_preds.device)
    _lib.call("crb3d_anchor_head_scores", _p(cls_preds), n, n_class, _p(score), _p(label), _stream(cls_preds.device))
    return score, label


def anchor_head_scores_topk(cls_preds, n_class, batch_size, thresh, k):
    """Fused score pass + candidate selection (class_agnostic_nms front half, model_nms_utils.py:6-25).
    cls_preds (B, A, n_class) channels-last logits. Returns score (B*A,), label (B*A,) int32 (1-based) for every anchor and,
    per frame, the anchors with score >= thresh sorted by descending score (ties: ascending anchor index), cut at k:
    top_scores (B, k) float32, top_idx (B, k) int64, counts (B,) int32 (valid prefix length; the tail is zero)."""
    _need_cuda(cls_preds)
    cls_preds = _f32c(cls_preds)
    dev = cls_preds.device
    B = int(batch_size)
    A = cls_preds.numel() // n_class // max(B, 1)
    score = torch.empty((B * A,), dtype=torch.float32, device=dev)
    label = torch.empty((B * A,), dtype=torch.int32, device=dev)
    cand = torch.empty((B, A), dtype=torch.int64, device=dev)
    cand_count = torch.empty((B * 2049,), dtype=torch.int32, device=dev)   # counts + 2048-bin score histogram per frame
    top_scores = torch.empty((B, k), dtype=torch.float32, device=dev)
    top_idx = torch.empty((B, k), dtype=torch.int64, device=dev)
    counts = torch.empty((B,), dtype=torch.int32, device=dev)
    _lib.call("crb3d_anchor_head_scores_topk", _p(cls_preds), B, A, n_class, float(thresh), int(k), _p(score), _p(label),
              _p(cand), _p(cand_count), _p(top_scores), _p(top_idx), _p(counts), _stream(dev))
    return score, label, top_scores, top_idx, counts


def anchor_decode_select(box_preds, dir_preds, sel, spec, n_anchor_per_frame):
    """Decode only the selected anchors. box_preds (B, A, 7), dir_preds (B, A, bins) | None, sel (B, K) int64."""
    _need_cuda(box_preds, sel)
    box_preds = _f32c(box_preds)
    dir_preds = _f32c(dir_preds) if dir_preds is not None else None
    sel = sel.contiguous()
    B, K = sel.shape
    out = torch.empty((B, K, 7), dtype=torch.float32, device=box_preds.device)
    _lib.call("crb3d_anchor_decode_select", _p(box_preds), _p(dir_preds), _p(sel), B, K, int(n_anchor_per_frame),
              ctypes.byref(spec), _p(out), _stream(box_preds.device))
    return out


def gather_rows(src, idx, valid=None, fill=0):
    """out[b,k,:] = src[b, idx[b,k], :] for k < valid[b] else fill. src (B, n, w) float32|int32, idx (B, K) int64."""
    _need_cuda(src, idx)
    src = src.contiguous()
    idx = idx.contiguous()
    B, n, w = src.shape
    K = idx.shape[1]
    out = torch.empty((B, K, w), dtype=src.dtype, device=src.device)
    if src.dtype == torch.float32:
        _lib.call("crb3d_gather_rows_f32", _p(src), _p(idx), _p(valid), B, K, n, w, float(fill), _p(out), _stream(src.device))
    elif src.dtype == torch.int32:
        _lib.call("crb3d_gather_rows_i32", _p(src), _p(idx), _p(valid), B, K, n, w, int(fill), _p(out), _stream(src.device))
    else:
        raise TypeError("gather_rows supports float32 / int32")
    return out
