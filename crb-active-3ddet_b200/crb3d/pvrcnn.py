"""PV-RCNN on the crb3d kernels.

The reference's PVRCNN detector (pcdet/models/detectors/pv_rcnn.py) already runs over this library through the drop-in
(spconv shim + pcdet.ops stand-ins: tests/test_gpu_reference_dropin.py). `accelerate(model)` additionally replaces, on the
inference path (eval mode, autograd off), the two places where the reference's module code - not a compiled op - is the
hot spot:

* every StackSAModuleMSG (pcdet/ops/pointnet2/pointnet2_stack/pointnet2_modules.py:30-112) of VoxelSetAbstraction
  (voxel_set_abstraction.py:334-411: SA_rawpoints + SA_layers) and of PVRCNNHead.roi_grid_pool (pvrcnn_head.py:68-114):
  per scale ONE ball query + ONE fused kernel (crb3d_sa_group_mlp_maxpool: group + shared MLP + BatchNorm + ReLU + max-pool),
  written straight into the scale's column slice of the concatenated output - the grouped (M, C+3, nsample) tensors and the
  MLP activations are never materialised;
* PVRCNNHead.shared_fc_layer (pvrcnn_head.py:21-33): its first layer, Conv1d(216*128 = 27 648 -> 256), is a split-K tcgen05
  GEMM (crb3d_fc_gemm_tf32) with the eval BatchNorm folded into the epilogue, computed ONCE per forward: the reference
  re-runs the whole stack in each of its SAMPLING_ROUND = 5 Monte-Carlo dropout rounds (pvrcnn_head.py:187-196) although the
  dropout sits behind that layer, so its input and output are the same in every round.

The module works on duck-typed instances (same attribute names as the reference classes); state_dict keys, training and
autograd are untouched - with autograd enabled or in train mode every module runs the reference's own forward.

`roi_head_gradient_embedding` is CRB stage 2 for detectors with a RoI head (crb_sampling.py:165-207): the gradient of the
RoI-head loss w.r.t. shared_fc_layer[4].weight under the hypothetical labels of stage 1, obtained with
torch.autograd.grad(loss, weight): identical to `loss.backward(); weight.grad` but nothing below that layer (VSA, BEV,
sparse backbone) is back-propagated.
"""
import types

import torch
import torch.nn as nn

from . import ops


# --------------------------------------------------------------------------------------------- fused SA module
def _fold_mlp(seq):
    """nn.Sequential(Conv2d 1x1 (no bias), BatchNorm2d, ReLU, ...) -> (widths, packed): per layer the transposed weight
    [C_in][C_out] with the eval-BatchNorm scale folded in, followed by the shift [C_out] (layout of csrc/sa_mlp.cu)."""
    mods = list(seq)
    widths, parts = None, []
    i = 0
    while i < len(mods):
        conv, bn = mods[i], mods[i + 1]
        assert isinstance(conv, nn.Conv2d) and conv.kernel_size == (1, 1) and isinstance(bn, nn.BatchNorm2d) and isinstance(mods[i + 2], nn.ReLU)
        w = conv.weight.detach().float().view(conv.out_channels, conv.in_channels)
        scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias.detach() - bn.running_mean * scale
        if conv.bias is not None:
            shift = shift + scale * conv.bias.detach()
        if widths is None:
            widths = [conv.in_channels]
        widths.append(conv.out_channels)
        parts += [(w * scale.view(-1, 1)).t().contiguous().view(-1), shift.float().contiguous()]
        i += 3
    return widths, torch.cat(parts).contiguous()


def _state_key(module):
    return tuple((t._version, t.data_ptr()) for t in list(module.parameters()) + list(module.buffers()))


def accelerate_sa_module(sa):
    """Gives a StackSAModuleMSG instance (attributes `groupers[k].radius / nsample`, `mlps[k]`, `pool_method`) the fused
    inference forward. The folded weights are rebuilt whenever a parameter / buffer changed (version-keyed cache)."""
    if getattr(sa, "_crb3d_fused", False) or getattr(sa, "pool_method", "max_pool") != "max_pool":
        return sa
    original = sa.forward

    def plan(self):
        key = _state_key(self)
        c = getattr(self, "_crb3d_plan", None)
        if c is None or c[0] != key:
            c = (key, [_fold_mlp(m) for m in self.mlps])
            self._crb3d_plan = c
        return c[1]

    def forward(self, xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features=None, empty_voxel_set_zeros=True):
        if self.training or torch.is_grad_enabled() or not xyz.is_cuda:
            return original(xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, features)
        layers = plan(self)
        M = new_xyz.shape[0]
        out = torch.empty((M, sum(w[-1] for w, _ in layers)), dtype=torch.float32, device=xyz.device)
        B = int(xyz_batch_cnt.numel())
        c0 = 0
        for g, (widths, packed) in zip(self.groupers, layers):
            idx = torch.empty((M, g.nsample), dtype=torch.int32, device=xyz.device)
            ops.ball_query(B, M, float(g.radius), int(g.nsample), new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx)
            ops.sa_group_mlp_maxpool(xyz, xyz_batch_cnt, features, new_xyz, new_xyz_batch_cnt, idx, widths, packed, out[:, c0:])
            c0 += widths[-1]
        return new_xyz, out

    sa.forward = types.MethodType(forward, sa)
    sa._crb3d_fused = True
    return sa


# --------------------------------------------------------------------------------------------- fused shared FC
class FusedSharedFC(nn.Sequential):
    """Drop-in for PVRCNNHead.shared_fc_layer (same children, same state_dict keys, `[4].weight` still addressable).
    Inference: layer 0 (Conv1d k=1 over the flattened RoI grid) + its BatchNorm + ReLU = one split-K tensor-core GEMM,
    cached across the Monte-Carlo rounds of one forward (same input tensor); the rest (Dropout, Conv1d 256->256, BatchNorm,
    ReLU) runs as the reference's modules, so the dropout masks are drawn exactly as in the reference."""

    def _first_layer(self, x):
        conv, bn = self[0], self[1]
        key = (_state_key(conv), _state_key(bn))
        c = getattr(self, "_crb3d_w", None)
        if c is None or c[0] != key:
            scale = bn.weight.detach() / torch.sqrt(bn.running_var + bn.eps)
            shift = bn.bias.detach() - bn.running_mean * scale
            if conv.bias is not None:
                shift = shift + scale * conv.bias.detach()
            w = ops.round_tf32(conv.weight.detach().float().view(conv.out_channels, conv.in_channels))
            c = (key, w, scale.float().contiguous(), shift.float().contiguous())
            self._crb3d_w = c
        _, w, scale, shift = c
        a = x.view(x.shape[0], x.shape[1])
        ck = (x.data_ptr(), x._version, tuple(x.shape))
        hit = getattr(self, "_crb3d_out", None)
        if hit is None or hit[0] != ck:
            hit = (ck, ops.fc_gemm(a, w, scale, shift, relu=True), x)      # x kept alive: the address stays unique
            self._crb3d_out = hit
        return hit[1].view(x.shape[0], -1, 1)

    def forward(self, x):
        conv = self[0]
        fused = (not self.training and not torch.is_grad_enabled() and x.is_cuda and isinstance(conv, nn.Conv1d)
                 and conv.kernel_size == (1,) and isinstance(self[1], nn.BatchNorm1d) and isinstance(self[2], nn.ReLU)
                 and conv.in_channels % 32 == 0 and conv.out_channels % 128 == 0 and x.shape[-1] == 1 and x.is_contiguous())
        if not fused:
            self._crb3d_out = None
            return super().forward(x)
        h = self._first_layer(x)
        for m in list(self)[3:]:
            h = m(h)
        return h


def accelerate_keypoint_sampling(pfe):
    """VoxelSetAbstraction.get_sampled_points (voxel_set_abstraction.py:227-283) samples the frames one after the other - 2048
    dependent FPS rounds each. With POINT_SOURCE = raw_points and SAMPLE_METHOD = FPS the frames are independent: ONE launch of
    the stacked FPS kernel (one thread-block cluster per frame) samples them side by side. Same indices: for clouds of at
    least 1024 points the stacked and the per-frame kernels of the reference share the 1024-thread tie rule (sampling_gpu.cu)."""
    cfg = getattr(pfe, "model_cfg", None)
    if cfg is None or getattr(pfe, "_crb3d_fps", False):
        return pfe
    get = (lambda k, d=None: cfg.get(k, d)) if hasattr(cfg, "get") else (lambda k, d=None: getattr(cfg, k, d))
    if get("POINT_SOURCE") != "raw_points" or get("SAMPLE_METHOD") != "FPS":
        return pfe
    original = pfe.get_sampled_points
    K = int(get("NUM_KEYPOINTS"))

    def get_sampled_points(self, batch_dict):
        pts = batch_dict["points"]
        B = int(batch_dict["batch_size"])
        if not pts.is_cuda or torch.is_grad_enabled():
            return original(batch_dict)
        bidx = pts[:, 0].long()
        cnt = torch.bincount(bidx, minlength=B).int()
        n_min, n_max = int(cnt.min()), int(cnt.max())          # one host read (the reference does several per frame)
        sorted_ok = bool((bidx[1:] >= bidx[:-1]).all()) if pts.shape[0] > 1 else True
        if n_min < max(K, 1024) or not sorted_ok:
            return original(batch_dict)
        xyz = pts[:, 1:4].contiguous()
        temp = torch.full((xyz.shape[0],), 1e10, dtype=torch.float32, device=pts.device)
        idx = torch.empty((B * K,), dtype=torch.int32, device=pts.device)
        ops.stack_farthest_point_sampling(xyz, temp, cnt, idx, torch.full((B,), K, dtype=torch.int32, device=pts.device), n_max=n_max)
        kp = xyz[idx.long()]
        b = torch.arange(B, device=pts.device, dtype=torch.float32).view(-1, 1).repeat(1, K).view(-1, 1)
        return torch.cat((b, kp), dim=1)

    pfe.get_sampled_points = types.MethodType(get_sampled_points, pfe)
    pfe._crb3d_fps = True
    return pfe


def accelerate(model, bev=True):
    """Patches a PV-RCNN detector instance in place (see the module docstring) and returns it. bev=True also applies
    crb3d.dropin.accelerate_bev_backbone (TF32 tensor-core plan) to its BEV backbone when that has the reference structure."""
    pfe = getattr(model, "pfe", None)
    if pfe is not None:
        accelerate_keypoint_sampling(pfe)
        if hasattr(pfe, "SA_rawpoints"):
            accelerate_sa_module(pfe.SA_rawpoints)
        for sa in getattr(pfe, "SA_layers", []):
            accelerate_sa_module(sa)
    head = getattr(model, "roi_head", None)
    if head is not None:
        if hasattr(head, "roi_grid_pool_layer"):
            accelerate_sa_module(head.roi_grid_pool_layer)
        fc = getattr(head, "shared_fc_layer", None)
        if isinstance(fc, nn.Sequential) and not isinstance(fc, FusedSharedFC):
            head.shared_fc_layer = FusedSharedFC(*list(fc))
            head.shared_fc_layer.training = fc.training      # the container's flag only: the children keep their own modes
            #                                                  (CRB switches the Dropout layers to train() for its MC rounds)
    b2d = getattr(model, "backbone_2d", None)
    if bev and b2d is not None and hasattr(b2d, "blocks") and hasattr(b2d, "deblocks") and not hasattr(b2d, "forward_inference"):
        from . import dropin
        try:
            dropin.accelerate_bev_backbone(b2d)
        except Exception:
            pass
    return model


# --------------------------------------------------------------------------------------------- CRB stage 2 (RoI head)
def roi_head_gradient_embedding(model, batch_dict, rcnn_cls_labels, rcnn_reg_targets):
    """crb_sampling.py:187-207 for ONE frame: train-mode forward of `model` on `batch_dict`, RoI-head classification +
    regression losses against the hypothetical labels (stage-1 Monte-Carlo means), and the gradient w.r.t.
    roi_head.shared_fc_layer[4].weight, flattened (256*256,). Uses torch.autograd.grad: the same tensor the reference reads
    from `.weight.grad` after `loss.backward()`, without back-propagating below that layer."""
    head = model.roi_head
    ret = model(batch_dict)
    pred = ret[0] if isinstance(ret, (tuple, list)) else ret
    cls_loss, _ = head.get_box_cls_layer_loss({"rcnn_cls": pred["rcnn_cls"], "rcnn_cls_labels": rcnn_cls_labels})
    reg_loss = head.get_box_reg_layer_loss({"rcnn_reg": pred["rcnn_reg"], "reg_sample_targets": rcnn_reg_targets})
    reg_loss = reg_loss[0] if isinstance(reg_loss, (tuple, list)) else reg_loss
    loss = cls_loss + reg_loss.mean()
    (g,) = torch.autograd.grad(loss, head.shared_fc_layer[4].weight, retain_graph=False, allow_unused=False)
    return g.detach().reshape(-1)
