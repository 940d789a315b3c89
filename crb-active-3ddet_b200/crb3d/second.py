"""SECOND detector + CRB stage-1 scoring on the crb3d kernels.

Module and parameter names mirror the reference so its checkpoints load unchanged
(pcdet/models/detectors/second_net.py:9-22, detector3d_template.py:24-53 module_topology):
  vfe (MeanVFE, fused into the voxelizer) -> backbone_3d (VoxelBackBone8x, spconv_backbone.py:69-180) ->
  map_to_bev_module (HeightCompression) -> backbone_2d (BaseBEVBackbone, base_bev_backbone.py:7-112) ->
  dense_head (AnchorHeadSingle, anchor_head_single.py:6-76) -> post_processing (detector3d_template.py:186-409).
`score_batch` is the pool-scoring hot path (forward + per-frame CRB stage-1 record) with no host synchronisation after
the rulebook phase: max-class scores, top-k, lazy decode, batched rotated NMS, points-in-boxes density and the label
entropy all stay on the device.
"""
import os
from functools import partial

import numpy as np
import torch
import torch.nn as nn

import spconv.pytorch as spconv

from . import head_ops, ops, synth

KITTI_SECOND_CFG = dict(
    class_names=["Car", "Pedestrian", "Cyclist"],
    data=synth.KITTI,
    num_bev_features=256,
    layer_nums=[5, 5], layer_strides=[1, 2], num_filters=[128, 256], upsample_strides=[1, 2], num_upsample_filters=[256, 256],
    anchors=[
        dict(class_name="Car", anchor_sizes=[[3.9, 1.6, 1.56]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[-1.78],
             matched_threshold=0.6, unmatched_threshold=0.45),
        dict(class_name="Pedestrian", anchor_sizes=[[0.8, 0.6, 1.73]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[-0.6],
             matched_threshold=0.5, unmatched_threshold=0.35),
        dict(class_name="Cyclist", anchor_sizes=[[1.76, 0.6, 1.73]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[-0.6],
             matched_threshold=0.5, unmatched_threshold=0.35),
    ],
    loss=dict(cls_weight=1.0, loc_weight=2.0, dir_weight=0.2, code_weights=[1.0] * 7),            # second.yaml:78-85
    dir_offset=0.78539, dir_limit_offset=0.0, num_dir_bins=2,
    score_thresh=0.1, nms_thresh=0.01, nms_pre_maxsize=4096, nms_post_maxsize=500,   # second.yaml:88-99
)

WAYMO_SECOND_CFG = dict(
    class_names=["Vehicle", "Pedestrian", "Cyclist"],
    data=synth.WAYMO,
    num_bev_features=256,
    layer_nums=[5, 5], layer_strides=[1, 2], num_filters=[128, 256], upsample_strides=[1, 2], num_upsample_filters=[256, 256],
    anchors=[  # tools/cfgs/waymo_models/second.yaml
        dict(class_name="Vehicle", anchor_sizes=[[4.7, 2.1, 1.7]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[0],
             matched_threshold=0.55, unmatched_threshold=0.4),
        dict(class_name="Pedestrian", anchor_sizes=[[0.91, 0.86, 1.73]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[0],
             matched_threshold=0.5, unmatched_threshold=0.35),
        dict(class_name="Cyclist", anchor_sizes=[[1.78, 0.84, 1.78]], anchor_rotations=[0, 1.57], anchor_bottom_heights=[0],
             matched_threshold=0.5, unmatched_threshold=0.35),
    ],
    loss=dict(cls_weight=1.0, loc_weight=2.0, dir_weight=0.2, code_weights=[1.0] * 7),
    dir_offset=0.78539, dir_limit_offset=0.0, num_dir_bins=2,
    score_thresh=0.1, nms_thresh=0.7, nms_pre_maxsize=4096, nms_post_maxsize=500,
)


# debug: one-thread marker kernels between the stages of the whole-step graph (tools/stress_hang.py reads them after a hang)
MARKERS = bool(int(os.environ.get("CRB3D_MARKERS", "0")))

# tcgen05 halo-tile kernel for the 3x3 stride-1 BEV convs: "auto" (when its CTA count fills whole waves), True, False
BEV_CONV_TC = {"0": False, "1": True}.get(os.environ.get("CRB3D_BEV_CONV_TC", ""), "auto")
# block 1 of the BEV backbone in sparse-tile mode (constant tiles far from every occupied cell are filled, not computed)
BEV_SPARSE_TILES = os.environ.get("CRB3D_BEV_SPARSE_TILES", "1") != "0"


_tf32_weight = ops.tf32_weight


def post_act_block(in_channels, out_channels, kernel_size, indice_key=None, stride=1, padding=0, conv_type="subm", norm_fn=None):
    if conv_type == "subm":
        conv = spconv.SubMConv3d(in_channels, out_channels, kernel_size, bias=False, indice_key=indice_key)
    elif conv_type == "spconv":
        conv = spconv.SparseConv3d(in_channels, out_channels, kernel_size, stride=stride, padding=padding, bias=False,
                                   indice_key=indice_key)
    elif conv_type == "inverseconv":
        conv = spconv.SparseInverseConv3d(in_channels, out_channels, kernel_size, indice_key=indice_key, bias=False)
    else:
        raise NotImplementedError
    return spconv.SparseSequential(conv, norm_fn(out_channels), nn.ReLU())


class VoxelBackBone8x(nn.Module):
    def __init__(self, input_channels, grid_size):
        super().__init__()
        norm_fn = partial(nn.BatchNorm1d, eps=1e-3, momentum=0.01)
        self.sparse_shape = [int(grid_size[2]) + 1, int(grid_size[1]), int(grid_size[0])]   # z gets +1 (spconv_backbone.py:75)
        self.conv_input = spconv.SparseSequential(
            spconv.SubMConv3d(input_channels, 16, 3, padding=1, bias=False, indice_key="subm1"), norm_fn(16), nn.ReLU())
        block = post_act_block
        self.conv1 = spconv.SparseSequential(block(16, 16, 3, norm_fn=norm_fn, padding=1, indice_key="subm1"))
        self.conv2 = spconv.SparseSequential(
            block(16, 32, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key="spconv2", conv_type="spconv"),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key="subm2"),
            block(32, 32, 3, norm_fn=norm_fn, padding=1, indice_key="subm2"))
        self.conv3 = spconv.SparseSequential(
            block(32, 64, 3, norm_fn=norm_fn, stride=2, padding=1, indice_key="spconv3", conv_type="spconv"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm3"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm3"))
        self.conv4 = spconv.SparseSequential(
            block(64, 64, 3, norm_fn=norm_fn, stride=2, padding=(0, 1, 1), indice_key="spconv4", conv_type="spconv"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm4"),
            block(64, 64, 3, norm_fn=norm_fn, padding=1, indice_key="subm4"))
        self.conv_out = spconv.SparseSequential(
            spconv.SparseConv3d(64, 128, (3, 1, 1), stride=(2, 1, 1), padding=0, bias=False, indice_key="spconv_down2"),
            norm_fn(128), nn.ReLU())
        self.num_point_features = 128

    def build_rulebooks(self, coords, batch_size):
        """Geometry pass: all 8 rulebooks (4 SubM keys, 4 strided) from the voxel coordinates alone, before any feature
        kernel is queued. The 4 host reads of n_out then only wait for tiny rulebook kernels, and the 12 conv launches
        that follow run back to back."""
        g = spconv.SparseConvTensor(None, coords, self.sparse_shape, batch_size)
        for seq in (self.conv_input, self.conv1, self.conv2, self.conv3, self.conv4, self.conv_out):
            for m in seq.modules():
                if isinstance(m, spconv.SparseConvolution):
                    d = m._rulebook(g)
                    if not m.subm:
                        g = spconv.SparseConvTensor(None, d.out_indices, d.out_spatial_shape, batch_size, indice_dict=g.indice_dict)
        return g.indice_dict

    def forward(self, batch_dict):
        coords = batch_dict["voxel_coords"].int()
        books = batch_dict.get("rulebooks")
        if books is None:
            books = self.build_rulebooks(coords, batch_dict["batch_size"])
        x = spconv.SparseConvTensor(features=batch_dict["voxel_features"], indices=coords,
                                    spatial_shape=self.sparse_shape, batch_size=batch_dict["batch_size"], indice_dict=books)
        x = self.conv_input(x)
        x1 = self.conv1(x)
        x2 = self.conv2(x1)
        x3 = self.conv3(x2)
        x4 = self.conv4(x3)
        out = self.conv_out(x4)
        batch_dict.update({"encoded_spconv_tensor": out, "encoded_spconv_tensor_stride": 8,
                           "multi_scale_3d_features": {"x_conv1": x1, "x_conv2": x2, "x_conv3": x3, "x_conv4": x4},
                           "multi_scale_3d_strides": {"x_conv1": 1, "x_conv2": 2, "x_conv3": 4, "x_conv4": 8}})
        return batch_dict


class HeightCompression(nn.Module):
    def __init__(self, num_bev_features):
        super().__init__()
        self.num_bev_features = num_bev_features

    def forward(self, batch_dict):
        t = batch_dict["encoded_spconv_tensor"]
        # (B, C*D, H, W) values identical to dense().view(N, C*D, H, W); memory is channels-last for the BEV convs
        batch_dict["spatial_features"] = t.dense_bev_channels_last()
        batch_dict["spatial_features_stride"] = batch_dict["encoded_spconv_tensor_stride"]
        return batch_dict


class BaseBEVBackbone(nn.Module):
    def __init__(self, cfg, input_channels):
        super().__init__()
        layer_nums, layer_strides, num_filters = cfg["layer_nums"], cfg["layer_strides"], cfg["num_filters"]
        ups, num_up = cfg["upsample_strides"], cfg["num_upsample_filters"]
        c_in_list = [input_channels, *num_filters[:-1]]
        self.blocks, self.deblocks = nn.ModuleList(), nn.ModuleList()
        for idx in range(len(layer_nums)):
            layers = [nn.ZeroPad2d(1), nn.Conv2d(c_in_list[idx], num_filters[idx], 3, stride=layer_strides[idx], padding=0, bias=False),
                      nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()]
            for _ in range(layer_nums[idx]):
                layers += [nn.Conv2d(num_filters[idx], num_filters[idx], 3, padding=1, bias=False),
                           nn.BatchNorm2d(num_filters[idx], eps=1e-3, momentum=0.01), nn.ReLU()]
            self.blocks.append(nn.Sequential(*layers))
            self.deblocks.append(nn.Sequential(
                nn.ConvTranspose2d(num_filters[idx], num_up[idx], ups[idx], stride=ups[idx], bias=False),
                nn.BatchNorm2d(num_up[idx], eps=1e-3, momentum=0.01), nn.ReLU()))
        self.num_bev_features = sum(num_up)

    # -- inference plan: eval-mode BatchNorm folded into the conv weights, ReLU fused into the conv call --------------
    @staticmethod
    def _fold(w, bn, transposed=False):
        scale = bn.weight / torch.sqrt(bn.running_var + bn.eps)
        shift = bn.bias - bn.running_mean * scale
        w = w * (scale.view(1, -1, 1, 1) if transposed else scale.view(-1, 1, 1, 1))
        return w.contiguous(memory_format=torch.channels_last), shift.contiguous()

    def build_inference_plan(self):
        """[(conv layers, deblock, deblock GEMM)] per block; rebuilt whenever parameters change (call after loading a
        checkpoint / moving the module). 3x3 stride-1 convs get a packed weight for crb3d's tcgen05 halo-tile conv
        (ops.bev_conv3x3); deblocks with kernel == stride in {1, 2} run on the tcgen05 GEMM (ops.bev_gemm) and write
        their channel slice of the concatenated map directly. Everything else stays on cuDNN."""
        plan = []
        with torch.no_grad():
            for blk, de in zip(self.blocks, self.deblocks):
                layers = []
                mods = list(blk)
                convs = [(mods[1], mods[2], (1, 1))] + [(mods[j], mods[j + 1], mods[j].padding) for j in range(4, len(mods), 3)]
                for conv, bn, pad in convs:                              # ZeroPad2d(1) + conv(pad 0) == conv(pad 1)
                    w, b = self._fold(conv.weight, bn)
                    wpack = w2 = None
                    if (conv.kernel_size == (3, 3) and conv.stride == (1, 1) and tuple(pad) == (1, 1) and w.shape[0] % 128 == 0
                            and w.shape[1] % 16 == 0 and w.is_cuda):
                        wpack = ops.pack_conv3x3_weight(w)
                    if (conv.kernel_size in ((3, 3), (1, 1)) and conv.stride in ((1, 1), (2, 2)) and pad[0] == pad[1]
                            and w.shape[0] in (128, 256) and w.shape[1] % 32 == 0 and w.is_cuda):
                        w2 = ops.pack_conv_gemm_weight(w)         # implicit GEMM over strided TMA boxes: the stride-2 layer
                    layers.append((w, b, conv.stride, tuple(pad), wpack, w2))
                dw, db = self._fold(de[0].weight, de[1], transposed=True)
                st = de[0].stride[0] if isinstance(de[0], nn.ConvTranspose2d) else 0
                cin, cout = dw.shape[0], dw.shape[1]
                gemm = None
                if (st in (1, 2) and de[0].stride == (st, st) and de[0].kernel_size == (st, st) and cin % 32 == 0
                        and cout in (128, 256) and dw.is_cuda):
                    # ConvTranspose2d weight (C_in, C_out, kh, kw) -> [(dy,dx)][C_out][C_in]
                    gemm = (ops.round_tf32(dw.permute(2, 3, 1, 0).reshape(st * st * cout, cin).contiguous()), db, st)
                plan.append((layers, (dw, db, de[0].stride), gemm))
        self._plan = plan
        self._sparse_fills = self._build_sparse_fills(plan)
        return plan

    @staticmethod
    def _build_sparse_fills(plan):
        """Constants of the sparse-tile mode (ops.bev_tile_plan) for the leading 3x3 stride-1 layers of block 1: the input of the
        block is zero away from the occupied cells, so far from them (and from the border) layer l outputs one constant vector
        c_{l+1} = layer_l(constant field c_l), c_0 = 0. Each constant is read off the kernel itself run on a constant image, i.e.
        it has exactly the bits the dense computation produces there. None when the first block does not start with such layers."""
        fills = []
        layers = plan[0][0]
        c = None
        for w, b, stride, pad, wpack, w2 in layers:
            if wpack is None or not w.is_cuda:
                break
            cin = w.shape[1]
            x = torch.zeros((1, 24, 48, cin), dtype=torch.float32, device=w.device) if c is None else c.view(1, 1, 1, cin).expand(1, 24, 48, cin).contiguous()
            y = ops.bev_conv3x3(x, wpack, b, True, round_out=True)
            c = y[0, 12, 24].clone()
            if not bool((y[0, 8:16, 16:32] == c).all()):          # the interior of a constant image must be constant
                break
            fills.append(c)
        return fills or None

    @staticmethod
    def _tc_conv_pays(B, H, W, cout, cin=128):
        """Halo-tile CTA-pair kernel (crb3d_bev_conv3x3_tf32) vs the implicit GEMM over strided TMA boxes
        (crb3d_bev_conv_gemm_tf32) for a 3x3 stride-1 layer both can take: the pair kernel is persistent over 74 CTA pairs
        per 128-channel slice and picks 1- or 2-tile work items itself, so it wins as soon as its (tile, slice) units give
        every CTA something to do; tiny maps go to the finer-grained implicit GEMM (BEV_CONV_TC = True/False overrides)."""
        if BEV_CONV_TC != "auto":
            return bool(BEV_CONV_TC)
        u, v = (H, W) if H % 8 == 0 or W % 8 != 0 else (W, H)
        tiles = B * -(-u // 8) * -(-v // 16)
        return tiles * max(1, cout // 128) >= 148

    def forward_inference(self, x, mark=None, occupancy=None):
        """occupancy = (coords (n,4) [b,z,y,x], n_dev | None) of the sparse tensor whose dense() is x: block 1 then runs in
        sparse-tile mode (only the 128-pixel tiles near an occupied cell or the border go through the tensor cores; bit-identical)."""
        B = x.shape[0]
        mark = mark or (lambda stage: None)
        li = 0
        fills = getattr(self, "_sparse_fills", None)
        tplan = None
        if occupancy is not None and fills and BEV_SPARSE_TILES and self._tc_conv_pays(B, x.shape[2], x.shape[3], 128):
            # a prebuilt plan (dict: the captured step builds it on its geometry branch, off the critical path) or (coords, n_dev)
            tplan = occupancy if isinstance(occupancy, dict) else ops.bev_tile_plan(occupancy[0], occupancy[1], B, x.shape[2], x.shape[3], len(fills))
            if tplan["lists"].shape[0] < len(fills) or tplan["n_tiles"] != ops.bev_conv3x3_num_tiles(B, x.shape[2], x.shape[3]):
                tplan = None
        gemm_ok = all(g is not None for _, _, g in self._plan)
        if gemm_ok:
            ctot = sum(g[0].shape[0] // (g[2] * g[2]) for _, _, g in self._plan)
            cat = None
        xh = x.permute(0, 2, 3, 1)                                   # (B, H, W, C): the memory order of channels-last
        xh = xh if xh.is_contiguous() else xh.contiguous()
        ups, c0 = [], 0
        for bi, (layers, (dw, db, ds), gemm) in enumerate(self._plan):
            for lj, (w, b, stride, pad, wpack, w2) in enumerate(layers):
                if wpack is not None and (w2 is None or self._tc_conv_pays(B, xh.shape[1], xh.shape[2], w.shape[0], w.shape[1])):
                    tiles = (tplan, lj, fills[lj]) if (tplan is not None and bi == 0 and lj < len(fills)) else None
                    xh = ops.bev_conv3x3(xh, wpack, b, True, round_out=True, tiles=tiles)   # the next layer reads TF32 exactly
                elif w2 is not None and BEV_CONV_TC is not False:
                    xh = ops.bev_conv_gemm(xh, w2, b, w.shape[2], stride[0], pad[0], True, round_out=True)
                else:
                    xh = torch.cudnn_convolution_relu(xh.permute(0, 3, 1, 2), w, b, stride, pad, (1, 1), 1).permute(0, 2, 3, 1)
                    xh = xh if xh.is_contiguous() else xh.contiguous()
                li += 1
                mark(200 + li)
            if not gemm_ok:
                ups.append(torch.relu_(torch.nn.functional.conv_transpose2d(xh.permute(0, 3, 1, 2), dw, db, stride=ds)))
                continue
            gw, gb, st = gemm
            cout = gw.shape[0] // (st * st)
            h, wd = xh.shape[1], xh.shape[2]
            if cat is None:   # (B, H_out, W_out, sum C) channels-last; every deblock lands on the same output grid
                cat = torch.empty((B, h * st, wd * st, ctot), dtype=torch.float32, device=x.device)
            assert cat.shape[1] == h * st and cat.shape[2] == wd * st
            ops.bev_gemm(xh.view(B * h * wd, xh.shape[3]), gw, gb, True, [(cat.view(-1, ctot)[:, c0:], 0, cout, ctot)],
                         n_sub=st * st, up=2 if st == 2 else 0, in_hw=(h, wd), round_out=True)
            mark(300 + li)
            c0 += cout
        if gemm_ok:
            return cat.permute(0, 3, 1, 2)
        return torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]

    def forward(self, batch_dict):
        x = batch_dict["spatial_features"]
        if not self.training and not torch.is_grad_enabled() and getattr(self, "_plan", None) is not None:
            batch_dict["spatial_features_2d"] = self.forward_inference(x, batch_dict.get("_mark"), batch_dict.get("_occupancy"))
            return batch_dict
        ups = []
        for i in range(len(self.blocks)):
            x = self.blocks[i](x)
            ups.append(self.deblocks[i](x))
        batch_dict["spatial_features_2d"] = torch.cat(ups, dim=1) if len(ups) > 1 else ups[0]
        return batch_dict


class AnchorHeadSingle(nn.Module):
    def __init__(self, cfg, input_channels, grid_size):
        super().__init__()
        self.cfg = cfg
        self.num_class = len(cfg["class_names"])
        self.n_loc = sum(len(a["anchor_sizes"]) * len(a["anchor_rotations"]) * len(a["anchor_bottom_heights"]) for a in cfg["anchors"])
        self.conv_cls = nn.Conv2d(input_channels, self.n_loc * self.num_class, 1)
        self.conv_box = nn.Conv2d(input_channels, self.n_loc * 7, 1)
        self.conv_dir_cls = nn.Conv2d(input_channels, self.n_loc * cfg["num_dir_bins"], 1)
        nn.init.constant_(self.conv_cls.bias, -np.log((1 - 0.01) / 0.01))
        nn.init.normal_(self.conv_box.weight, mean=0, std=0.001)
        fm = (int(grid_size[0]) // 8, int(grid_size[1]) // 8)   # feature_map_stride 8
        self.spec = head_ops.make_anchor_spec(cfg["anchors"], cfg["data"]["pc_range"], fm, cfg["dir_offset"],
                                              cfg["dir_limit_offset"], cfg["num_dir_bins"])
        self.num_anchors = fm[0] * fm[1] * self.n_loc

    # ---- training path (anchor_head_template.py:88-229): targets and losses on the device (csrc/train_ops.cu)
    def anchors_device(self, device):
        a = getattr(self, "_anchors_dev", None)
        if a is None or a.device != device:
            a = head_ops.anchors_tensor(self.spec).to(device)
            self._anchors_dev = a
        return a

    def assign_targets(self, gt_boxes):
        """gt_boxes (B, M, 8) -> dict(box_cls_labels (B,A) int32, box_reg_targets (B,A,7), reg_weights (B,A))
        (AnchorHeadTemplate.assign_targets -> AxisAlignedTargetAssigner.assign_targets)."""
        from . import train_ops
        ta = getattr(self, "target_assigner", None)
        if ta is None:
            ta = self.target_assigner = train_ops.AxisAlignedTargetAssigner(self.cfg["anchors"], self.cfg["class_names"])
        return ta.assign_targets(self.anchors_device(gt_boxes.device), gt_boxes)

    def get_loss(self):
        """rpn_loss = cls + loc + dir of the last training forward, and the reference's tb_dict (tensors, not .item() floats: no
        host synchronisation inside the training step)."""
        from . import train_ops
        r = self.forward_ret_dict
        cfg = dict(self.cfg["loss"], dir_offset=self.cfg["dir_offset"])
        losses = train_ops.anchor_head_loss(r["cls_preds"], r["box_preds"], r.get("dir_cls_preds"), r["box_cls_labels"],
                                            r["box_reg_targets"], self.anchors_device(r["cls_preds"].device), cfg)
        rpn_loss = losses.sum()
        return rpn_loss, {"rpn_loss_cls": losses[0].detach(), "rpn_loss_loc": losses[1].detach(), "rpn_loss_dir": losses[2].detach(),
                          "rpn_loss": rpn_loss.detach()}

    def build_inference_plan(self):
        """The three 1x1 head convs as one [n_cls + n_box + n_dir (padded to 80)][C_in] weight for ops.bev_gemm."""
        self._plan = None
        with torch.no_grad():
            ws = [m.weight.reshape(m.weight.shape[0], -1) for m in (self.conv_cls, self.conv_box, self.conv_dir_cls)]
            bs = [m.bias for m in (self.conv_cls, self.conv_box, self.conv_dir_cls)]
            n, k = sum(w.shape[0] for w in ws), ws[0].shape[1]
            if n <= 80 and k % 32 == 0:
                w = torch.zeros((80, k), dtype=torch.float32, device=ws[0].device)
                b = torch.zeros((80,), dtype=torch.float32, device=ws[0].device)
                w[:n] = torch.cat(ws, 0)
                b[:n] = torch.cat(bs, 0)
                self._plan = (ops.round_tf32(w), b, [x.shape[0] for x in ws])
        return self._plan

    def forward(self, batch_dict):
        x = batch_dict["spatial_features_2d"]
        B = x.shape[0]
        plan = getattr(self, "_plan", None)
        if plan is not None and not self.training and not torch.is_grad_enabled() and x.is_cuda:
            w, b, widths = plan
            a = x.permute(0, 2, 3, 1)
            a = (a if a.is_contiguous() else a.contiguous()).view(-1, x.shape[1])
            outs = [torch.empty((a.shape[0], n), dtype=torch.float32, device=x.device) for n in widths]
            segs, c0 = [], 0
            for o, n in zip(outs, widths):
                segs.append((o, c0, n, n))
                c0 += n
            ops.bev_gemm(a, w, b, False, segs)
            batch_dict["cls_preds"] = outs[0].view(B, self.num_anchors, self.num_class)
            batch_dict["box_preds"] = outs[1].view(B, self.num_anchors, 7)
            batch_dict["dir_cls_preds"] = outs[2].view(B, self.num_anchors, -1)
            return batch_dict
        # NHWC: with channels-last activations permute(0,2,3,1) is a view and .contiguous() is free
        batch_dict["cls_preds"] = self.conv_cls(x).permute(0, 2, 3, 1).contiguous().view(B, self.num_anchors, self.num_class)
        batch_dict["box_preds"] = self.conv_box(x).permute(0, 2, 3, 1).contiguous().view(B, self.num_anchors, 7)
        batch_dict["dir_cls_preds"] = self.conv_dir_cls(x).permute(0, 2, 3, 1).contiguous().view(B, self.num_anchors, -1)
        if self.training and "gt_boxes" in batch_dict:       # anchor_head_single.py:60-66
            self.forward_ret_dict = {k: batch_dict[k] for k in ("cls_preds", "box_preds", "dir_cls_preds")}
            self.forward_ret_dict.update(self.assign_targets(batch_dict["gt_boxes"]))
        return batch_dict


class SECONDNet(nn.Module):
    def __init__(self, cfg=KITTI_SECOND_CFG):
        super().__init__()
        self.cfg = cfg
        d = cfg["data"]
        rng = np.asarray(d["pc_range"], dtype=np.float64)
        self.grid_size = np.round((rng[3:6] - rng[0:3]) / np.asarray(d["voxel_size"], dtype=np.float64)).astype(np.int64)
        self.num_class = len(cfg["class_names"])
        self.backbone_3d = VoxelBackBone8x(d["n_feat"], self.grid_size)
        self.map_to_bev_module = HeightCompression(cfg["num_bev_features"])
        self.backbone_2d = BaseBEVBackbone(cfg, cfg["num_bev_features"])
        self.dense_head = AnchorHeadSingle(cfg, self.backbone_2d.num_bev_features, self.grid_size)

    def to_device(self, device):
        self.to(device)
        self.backbone_2d.to(memory_format=torch.channels_last)
        self.dense_head.to(memory_format=torch.channels_last)
        return self

    def prepare_inference(self, fold_bev_bn=True, spconv_tf32=None):
        """Eval-time plan: BEV BatchNorms folded into their convs + fused ReLU; optionally switches the sparse convs
        with C_in >= 16 to the tcgen05 TF32 kernel (None = leave the global crb3d.ops.SPCONV_TF32 setting alone)."""
        self.eval()
        self.backbone_2d._plan = None
        self.dense_head._plan = None
        if fold_bev_bn:
            self.backbone_2d.build_inference_plan()
            self.dense_head.build_inference_plan()
        if spconv_tf32 is not None:
            ops.SPCONV_TF32 = bool(spconv_tf32)
        self._inference_args = (fold_bev_bn, spconv_tf32)
        self._inference_key = self._state_key()
        return self

    # -- cached inference state (folded BEV / head weights, captured graphs) vs training ------------------------------
    def _state_key(self):
        """Version counter + address of every parameter and buffer: changes on optimizer steps, load_state_dict, BatchNorm
        running-statistic updates and device moves (the same key SparseSequential._bn_affine uses per layer)."""
        return tuple((t._version, t.data_ptr()) for t in list(self.parameters()) + list(self.buffers()))

    def _drop_inference_state(self):
        self.backbone_2d._plan = None
        self.dense_head._plan = None
        ops.drop_workspaces([id(g) for g in getattr(self, "_full_graphs", [])])
        for name in ("_graph", "_full_graph", "_full_graphs"):
            if hasattr(self, name):
                delattr(self, name)

    def train(self, mode=True):
        """Training invalidates the inference plan and every captured graph (they hold folded copies of the weights and
        raw pointers to the BatchNorm affine tensors): the next ensure_inference_current() rebuilds them."""
        if mode:
            self._drop_inference_state()
        return super().train(mode)

    def ensure_inference_current(self):
        """Rebuilds the inference plan and re-captures the whole-step graphs when any parameter / buffer changed since they
        were built (active learning alternates train -> query -> train: CRBSampling.query and PoolScorer call this before
        scoring). Returns True when something was rebuilt."""
        args = getattr(self, "_inference_args", None)
        if args is None:
            return False
        stale = getattr(self, "_inference_key", None) != self._state_key()
        missing = (args[0] and self.backbone_2d._plan is None) or \
                  (getattr(self, "_graph_cfg", None) is not None and not hasattr(self, "_full_graph"))
        if not stale and not missing:
            return False
        gcfg = getattr(self, "_graph_cfg", None)
        self._drop_inference_state()
        self.prepare_inference(*args)
        if gcfg is not None:
            self.enable_full_graph(*gcfg)
        return True

    # ------------------------------------------------------------------------------------------- forward pieces
    def voxelize(self, points, frame_offsets, batch_size, training=False):
        d = self.cfg["data"]
        mv = d["max_voxels_train"] if training else d["max_voxels_test"]
        return ops.voxelize(points, frame_offsets, batch_size, d["pc_range"], d["voxel_size"], d["max_pts"], mv,
                            xyz_col=points.shape[1] - d["n_feat"], feat_col=points.shape[1] - d["n_feat"], n_feat=d["n_feat"])

    def geometry(self, points, frame_offsets, batch_size):
        """Everything that needs a host-visible count (voxel count + 4 strided-conv output counts): voxelize + MeanVFE
        and the 8 rulebooks. Can run on a side stream one batch ahead of the feature phase (PoolScorer.score_stream)."""
        vox = self.voxelize(points, frame_offsets, batch_size)
        books = self.backbone_3d.build_rulebooks(vox["coords"], batch_size)
        return dict(voxel_features=vox["mean"], voxel_coords=vox["coords"], rulebooks=books)

    def forward_features(self, points, frame_offsets, batch_size, geom=None, gt_boxes=None):
        """points (N, C) or (N, 1+C) with the batch index in column 0 (collate layout). Returns the head outputs. In training
        mode with gt_boxes (B, M, 8) the head also assigns its targets (dense_head.get_loss() then gives rpn_loss)."""
        if geom is None:
            geom = self.geometry(points, frame_offsets, batch_size)
        bd = dict(batch_size=batch_size, **geom)
        if gt_boxes is not None:
            bd["gt_boxes"] = gt_boxes
        bd = self.backbone_3d(bd)
        bd = self.map_to_bev_module(bd)
        bd = self.backbone_2d(bd)
        bd = self.dense_head(bd)
        return bd

    @torch.no_grad()
    def dense_and_post(self, spatial_features, points_xyz, pt_begin, pt_end, batch_size, max_pts_per_frame, mark=None, occupancy=None):
        """Static-shape half of the step: BEV backbone -> anchor head -> max-class score / top-k / lazy decode -> batched
        rotated NMS -> points-in-boxes density -> label entropy. No host synchronisation and no data-dependent shape, so
        the whole thing is capturable in one CUDA graph (enable_cuda_graph)."""
        cfg = self.cfg
        mark = mark or (lambda stage: None)
        # occupancy = (coords, n_dev) of the encoded sparse tensor: lets block 1 of the BEV backbone skip its constant tiles
        bd = self.backbone_2d(dict(spatial_features=spatial_features, _mark=mark, _occupancy=occupancy))
        mark(40)
        bd = self.dense_head(bd)
        mark(41)
        B, A = batch_size, self.dense_head.num_anchors
        dev = spatial_features.device
        # class_agnostic_nms (model_nms_utils.py:6-25): score >= thresh, top-k(NMS_PRE_MAXSIZE) - valid entries are a prefix
        k = min(cfg["nms_pre_maxsize"], A)
        if k <= 4096:
            score, label, top_scores, top_idx, counts = head_ops.anchor_head_scores_topk(bd["cls_preds"], self.num_class, B,
                                                                                         cfg["score_thresh"], k)
            label = label.view(B, A, 1)
        else:
            score, label = head_ops.anchor_head_scores(bd["cls_preds"], self.num_class)
            score, label = score.view(B, A), label.view(B, A, 1)
            top_scores, top_idx = torch.topk(score, k, dim=1)
            counts = (top_scores >= cfg["score_thresh"]).sum(dim=1).int()
        mark(42)
        boxes = head_ops.anchor_decode_select(bd["box_preds"], bd["dir_cls_preds"], top_idx, self.dense_head.spec, A)
        P = cfg["nms_post_maxsize"]
        mark(43)
        keep, num = ops.nms_batched(boxes, counts, cfg["nms_thresh"], rotated=True, max_keep=P)
        mark(44)
        final_boxes = head_ops.gather_rows(boxes, keep, num, 0.0)
        final_scores = head_ops.gather_rows(top_scores.unsqueeze(-1).contiguous(), keep, num, 0.0).squeeze(-1)
        anchor_of_kept = head_ops.gather_rows(top_idx.int().unsqueeze(-1).contiguous(), keep, num, 0).squeeze(-1)
        final_labels = head_ops.gather_rows(label, anchor_of_kept.long(), num, 0).squeeze(-1)
        # per-box point density (detector3d_template.py:379-387) and label entropy (crb_sampling.py:86-100)
        box_begin = torch.arange(B, device=dev, dtype=torch.int32) * P
        _, cnt, dens = ops.points_in_boxes_ranges(points_xyz, pt_begin, pt_end, max_pts_per_frame, final_boxes, box_begin,
                                                  box_begin + num)
        entropy = ops.label_entropy_ranges(final_labels, box_begin, box_begin + num, self.num_class)
        valid = torch.arange(P, device=dev).view(1, P) < num.view(B, 1)
        dens = torch.where(valid, dens.view(B, P), torch.zeros((), device=dev))   # padded slots: 0, not 0/0
        return dict(entropy=entropy, num_boxes=num, labels=final_labels, density=dens, point_counts=cnt.view(B, P),
                    boxes=final_boxes, scores=final_scores, anchor_idx=anchor_of_kept)

    @torch.no_grad()
    def enable_cuda_graph(self, batch_size, max_points_per_frame=32768):
        """Captures dense_and_post for a fixed batch size into a CUDA graph (~150 kernel launches -> 1 graph launch; at
        batch 4 the eager step is host-launch bound). Inputs live in static buffers: the BEV map the dense() scatter
        writes into and a point buffer of batch_size*max_points_per_frame rows for the density kernel."""
        dev = next(self.parameters()).device
        d = self.cfg["data"]
        C, (D, H, W) = self.backbone_3d.num_point_features, self.bev_shape()
        g = dict(B=batch_size, cap=batch_size * max_points_per_frame, max_pts=max_points_per_frame)
        g["spatial"] = torch.zeros((batch_size, H, W, C * D), device=dev)
        g["points"] = torch.zeros((g["cap"], d["n_feat"]), device=dev)
        g["offsets"] = torch.zeros((batch_size + 1,), dtype=torch.int32, device=dev)

        def body():
            return self.dense_and_post(g["spatial"].permute(0, 3, 1, 2), g["points"], g["offsets"][:-1], g["offsets"][1:],
                                       batch_size, max_points_per_frame)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(side):          # warm-up outside capture (cuDNN autotune, workspace allocations)
            for _ in range(3):
                body()
        torch.cuda.current_stream(dev).wait_stream(side)
        torch.cuda.synchronize(dev)
        graph = torch.cuda.CUDAGraph()
        with torch.cuda.graph(graph):
            g["out"] = body()
        g["graph"] = graph
        self._graph = g
        return self

    # ------------------------------------------------------------------------------------------- whole-step CUDA graph
    def _sparse_layers(self):
        """[(conv, bn, relu)] of the 3-D backbone in execution order (spconv_backbone.py:77-117)."""
        layers = []
        bb = self.backbone_3d
        for seq in (bb.conv_input, bb.conv1, bb.conv2, bb.conv3, bb.conv4, bb.conv_out):
            for m in seq.modules():
                if isinstance(m, spconv.SparseConvolution):
                    layers.append([m, None, False])
                elif isinstance(m, nn.BatchNorm1d):
                    layers[-1][1] = m
                elif isinstance(m, nn.ReLU):
                    layers[-1][2] = True
        return layers

    def _static_step(self, g):
        """One whole scoring step on capacity-sized static buffers with DEVICE-side row counts: voxelize + MeanVFE, the 8
        rulebooks, 12 sparse convs (BN + ReLU fused), dense(), BEV stack, head, top-k, NMS, density, entropy. No host
        synchronisation anywhere, so the step is one CUDA graph (enable_full_graph). Counts that exceed their capacity
        are reported in `counts` (true values) next to `caps`; the caller re-scores such a batch on the dynamic path."""
        B = g["B"]
        d = self.cfg["data"]
        ops.WS_TAG = id(g)       # scratch buffers private to this graph copy (see ops._ws); reset by the caller
        mark = (lambda stage: ops.debug_mark(g.get("slot", 0) * 2, stage)) if MARKERS else (lambda stage: None)
        mark_side = (lambda stage: ops.debug_mark(g.get("slot", 0) * 2 + 1, stage)) if MARKERS else (lambda stage: None)
        mark(1)
        vox = ops.voxelize(g["points"], g["offsets"], B, d["pc_range"], d["voxel_size"], d["max_pts"], d["max_voxels_test"],
                           xyz_col=0, feat_col=0, n_feat=d["n_feat"], sync=False)
        feat, coords, n_dev = vox["mean"], vox["coords"], vox["n_dev"]
        mark(2)
        shape = list(self.backbone_3d.sparse_shape)
        # geometry (8 rulebooks: dozens of small latency-bound kernels) runs on a forked branch of the graph, concurrently
        # with the sparse convs of the earlier layers; each conv waits only for the event of ITS rulebook
        main = torch.cuda.current_stream(g["points"].device)
        side = g.setdefault("side", torch.cuda.Stream(g["points"].device))
        fork = torch.cuda.Event()
        fork.record(main)
        side.wait_event(fork)
        books, counts, caps, plan = {}, [n_dev], [coords.shape[0]], []
        cellmap = None          # cell -> row map of the current level (left behind by the strided rulebook that produced it)
        g["_cellmaps"] = []     # keeps those workspaces alive for as long as the captured graph
        with torch.cuda.stream(side):
            for conv, bn, relu in self._sparse_layers():
                mark_side(100 + len(plan))    # before this layer's rulebook (its event is recorded after, so the branch stays joined)
                if conv.subm:
                    key = (conv.indice_key, tuple(conv.kernel_size))
                    if key not in books:
                        nbr = ops.subm_rulebook(coords, shape, conv.kernel_size, conv.dilation, n_dev=n_dev, cellmap=cellmap)
                        ev = torch.cuda.Event()
                        ev.record(side)
                        books[key] = (nbr, ev)
                    nbr, ev = books[key]
                else:
                    cells = B * np.prod(ops.conv_out_shape(shape, conv.kernel_size, conv.stride, conv.padding, conv.dilation))
                    if g.get("row_caps") is not None:        # measured capacities (calibrate_row_caps): grids sized for the data
                        cap_out = int(min(cells, g["row_caps"][len(caps) - 1]))
                    else:
                        cap_out = int(min(cells, g["growth"][len(caps) - 1] * coords.shape[0]))
                    coords, shape, nbr, n_dev, cellmap = ops.sparse_rulebook_static(coords, n_dev, B, shape, conv.kernel_size, conv.stride,
                                                                                    conv.padding, cap_out, conv.dilation, want_cellmap=True)
                    g["_cellmaps"].append(cellmap)
                    ev = torch.cuda.Event()
                    ev.record(side)
                    counts.append(n_dev)
                    caps.append(cap_out)
                plan.append((conv, bn, relu, nbr, n_dev, ev))
            # sparse-tile plan of BEV block 1 from the encoded tensor's rows: known as soon as the last rulebook is, so it is built
            # here on the geometry branch, under the sparse convs
            occupancy = (coords, n_dev)
            fills = getattr(self.backbone_2d, "_sparse_fills", None)
            if fills and BEV_SPARSE_TILES:
                D2, H2, W2 = shape
                occupancy = ops.bev_tile_plan(coords, n_dev, B, H2, W2, len(fills))
                ev_occ = torch.cuda.Event()
                ev_occ.record(side)
                g["_tile_plan"] = occupancy
        for li, (conv, bn, relu, nbr, nd, ev) in enumerate(plan):
            main.wait_event(ev)
            scale, shift = spconv.SparseSequential._bn_affine(bn) if bn is not None else (None, None)
            if conv.bias is not None:
                shift = conv.bias if shift is None else shift + scale * conv.bias
            tc = ops.SPCONV_TF32
            feat = ops.spconv_forward(feat, nbr, _tf32_weight(conv) if tc else conv.weight, scale=scale, shift=shift, relu=relu,
                                      n_dev=nd, round_out=tc)
            mark(10 + li)
        n_dev = plan[-1][4]
        ops.sparse_to_dense(feat, coords, B, shape, channels_last_bev=True, out=g["spatial"], n_dev=n_dev)
        mark(30)
        if isinstance(occupancy, dict):
            main.wait_event(ev_occ)
        out = self.dense_and_post(g["spatial"].permute(0, 3, 1, 2), g["points"], g["offsets"][:-1], g["offsets"][1:], B,
                                  g["max_pts"], mark=mark, occupancy=occupancy)
        mark(99)
        out["counts"] = torch.cat(counts)
        g["caps"] = caps
        return out

    @torch.no_grad()
    @torch.no_grad()
    def level_row_counts(self, points, frame_offsets, batch_size):
        """Row counts of one batch at every level of the sparse backbone: [voxels, rows after each strided conv] (host ints)."""
        geom = self.geometry(points, frame_offsets, batch_size)
        counts = [int(geom["voxel_coords"].shape[0])]
        seen = set()
        for conv, _, _ in self._sparse_layers():
            if not conv.subm and conv.indice_key not in seen:
                seen.add(conv.indice_key)
                counts.append(int(geom["rulebooks"][conv.indice_key].out_indices.shape[0]))
        return counts

    def calibrate_row_caps(self, batches, margin=1.3, slack=2048):
        """Static row capacities of the strided-conv outputs for enable_full_graph(row_caps=...), measured on sample batches
        [(points, frame_offsets, batch_size)]: max count per level x margin (+slack), rounded up to 128-row tiles. Every kernel
        of the captured step launches a capacity-sized grid (rows beyond the device-side count exit at once), so capacities
        sized for the data instead of the worst case (growth=2,1,1,1: every level as large as the first strided output) remove
        most CTAs of the deep layers. A batch that exceeds a capacity is detected from `counts` and re-scored eagerly."""
        mx = None
        for p, o, b in batches:
            c = self.level_row_counts(p, o, b)
            mx = c if mx is None else [max(a, x) for a, x in zip(mx, c)]
        return [int(-(-(int(c * margin) + slack) // 128) * 128) for c in mx[1:]]

    def enable_full_graph(self, batch_size, max_points_per_frame=32768, growth=(2.0, 1.0, 1.0, 1.0), slots=1, row_caps=None):
        """Captures the WHOLE scoring step (_static_step) for a fixed batch size into one CUDA graph: every count stays on
        the device, so one replay replaces ~250 kernel launches and 5 host synchronisations (the eager step is host-bound).
        growth[i]: capacity of the i-th strided conv's output rows relative to its input capacity; row_caps (optional, wins):
        absolute capacities of the strided-conv outputs, e.g. from calibrate_row_caps().
        slots > 1 captures that many independent copies (own static buffers): replayed on different streams, consecutive
        batches overlap - the narrow tail of one step (greedy NMS pass, top-k sort, small rulebook kernels, ~10 % of the
        step on a handful of SMs) runs under the wide kernels of the next one."""
        row_caps = None if row_caps is None else tuple(int(c) for c in row_caps)
        self._graph_cfg = (batch_size, max_points_per_frame, tuple(growth), slots, row_caps)
        if slots > 1:
            ops.drop_workspaces([id(g) for g in getattr(self, "_full_graphs", [])])     # re-capture: the old copies go away
            copies = []
            for i in range(slots):
                self._next_slot = i
                self.enable_full_graph(batch_size, max_points_per_frame, growth, slots=1, row_caps=row_caps)
                copies.append(self._full_graph)
            self._next_slot = 0
            self._full_graphs = copies
            self._full_graph = copies[0]
            self._graph_cfg = (batch_size, max_points_per_frame, tuple(growth), slots, row_caps)
            return self
        dev = next(self.parameters()).device
        d = self.cfg["data"]
        C, (D, H, W) = self.backbone_3d.num_point_features, self.bev_shape()
        g = dict(B=batch_size, cap=batch_size * max_points_per_frame, max_pts=max_points_per_frame, growth=list(growth),
                 row_caps=row_caps, slot=getattr(self, "_next_slot", 0))
        g["spatial"] = torch.zeros((batch_size, H, W, C * D), device=dev)
        g["points"] = torch.zeros((g["cap"], d["n_feat"]), device=dev)
        g["offsets"] = torch.zeros((batch_size + 1,), dtype=torch.int32, device=dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(torch.cuda.current_stream(dev))
        try:
            with torch.cuda.stream(side):          # warm-up outside capture (cuDNN autotune, workspace allocations)
                for _ in range(3):
                    self._static_step(g)
            torch.cuda.current_stream(dev).wait_stream(side)
            torch.cuda.synchronize(dev)
            from . import _lib
            k0 = _lib.LAUNCHES["kernels"]
            graph = torch.cuda.CUDAGraph()
            with torch.cuda.graph(graph):
                g["out"] = self._static_step(g)
        finally:
            ops.WS_TAG = None
        g["kernels"] = _lib.LAUNCHES["kernels"] - k0      # kernels of libcrb3d_sm100 recorded in the graph (per replay)
        g["graph"] = graph
        self._full_graph = g
        self._full_graphs = [g]
        return self

    def full_graph_replay(self, points, frame_offsets, slot=0):
        """Copies one batch (device or pinned host tensors) into the static buffers of graph copy `slot` and replays it on
        the current stream. Returns the static output dict (consume or copy it on the same stream before the next replay
        of that slot); out["counts"] vs self._full_graph["caps"] tells whether a capacity was exceeded."""
        g = self._full_graphs[slot]
        n = points.shape[0]
        if n > g["cap"] or frame_offsets.numel() != g["B"] + 1:
            raise ValueError("batch does not fit the captured graph")
        xyz_col = points.shape[1] - self.cfg["data"]["n_feat"]
        g["points"][:n].copy_(points[:, xyz_col:], non_blocking=True)
        g["offsets"].copy_(frame_offsets, non_blocking=True)
        g["graph"].replay()
        from . import _lib
        _lib.LAUNCHES["kernels"] += g["kernels"]
        return g["out"]

    def bev_shape(self):
        """(D, H, W) of the encoded sparse tensor (z: 41 -> 21 -> 11 -> 5 -> 2; y, x: /8)."""
        s = list(self.backbone_3d.sparse_shape)
        for k, st, p in (((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                         ((3, 1, 1), (2, 1, 1), (0, 0, 0))):
            s = ops.conv_out_shape(s, k, st, p)
        return s

    @torch.no_grad()
    def score_batch(self, points, frame_offsets, batch_size, max_pts_per_frame, geom=None):
        """CRB stage-1 record of every frame of the batch (crb_sampling.py:72-103 via post_processing):
        dict(entropy (B,), num_boxes (B,), labels (B,P) int32 1-based (0 pad), density (B,P), boxes (B,P,7), scores (B,P)).
        With enable_cuda_graph() and a matching batch the returned tensors are the graph's static outputs: consume (or
        copy) them on the same stream before the next call."""
        fg = getattr(self, "_full_graph", None)
        if (geom is None and fg is not None and fg["B"] == batch_size and points.shape[0] <= fg["cap"]
                and max_pts_per_frame <= fg["max_pts"]):
            return self.full_graph_replay(points, frame_offsets)
        if geom is None:
            geom = self.geometry(points, frame_offsets, batch_size)
        bd = self.backbone_3d(dict(batch_size=batch_size, **geom))
        enc = bd["encoded_spconv_tensor"]
        xyz_col = points.shape[1] - self.cfg["data"]["n_feat"]
        g = getattr(self, "_graph", None)
        if g is not None and g["B"] == batch_size and points.shape[0] <= g["cap"] and max_pts_per_frame <= g["max_pts"]:
            ops.sparse_to_dense(enc.features, enc.indices, batch_size, enc.spatial_shape, channels_last_bev=True, out=g["spatial"])
            g["points"][: points.shape[0]].copy_(points[:, xyz_col:], non_blocking=True)
            g["offsets"].copy_(frame_offsets, non_blocking=True)
            g["graph"].replay()
            return g["out"]
        spatial = enc.dense_bev_channels_last()
        return self.dense_and_post(spatial, points[:, xyz_col:], frame_offsets[:-1], frame_offsets[1:], batch_size,
                                   max_pts_per_frame, occupancy=(enc.indices, None))


def calibrate_batchnorm(model, points, frame_offsets, batch_size):
    """Synthetic weights only: sets every BatchNorm's running statistics to the statistics of one batch (one train-mode
    forward with momentum 1). With the default init (running_mean 0, running_var 1) the activations of the 24 randomly
    initialised layers decay geometrically (rms 1e-1 after conv1, 1e-10 at the head input on KITTI-shaped clouds), every
    class logit equals its bias and all 211 200 anchors tie at one score - a degenerate workload for top-k and NMS. A
    trained checkpoint carries real statistics; this gives the random model the same property."""
    bns = [m for m in model.modules() if isinstance(m, (nn.BatchNorm1d, nn.BatchNorm2d))]
    old = [m.momentum for m in bns]
    was_training = model.training
    plan2d, planh = getattr(model.backbone_2d, "_plan", None), getattr(model.dense_head, "_plan", None)
    model.backbone_2d._plan = None
    model.dense_head._plan = None
    try:
        model.train()
        for m in bns:
            m.momentum = 1.0
        with torch.no_grad():
            model.forward_features(points, frame_offsets, batch_size)
    finally:
        for m, mo in zip(bns, old):
            m.momentum = mo
        model.train(was_training)
    if plan2d is not None:
        model.backbone_2d.build_inference_plan()
    if planh is not None:
        model.dense_head.build_inference_plan()
    if getattr(model, "_inference_args", None) is not None:
        model._inference_key = model._state_key()      # the plans above were built from the calibrated statistics
    return model


def calibrate_head_bias(model, points, frame_offsets, batch_size, target_fraction=0.03):
    """Synthetic weights only: the default conv_cls init (bias -log(99), anchor_head_single.py:37-38) yields no box
    above SCORE_THRESH, and with random features one class's logits dominate every anchor. Each class's logits are
    standardised (conv_cls rows rescaled, bias shifted) to one common distribution whose upper `target_fraction` tail
    clears the threshold, so all classes get predicted (CRB stage 3 needs every class)."""
    with torch.no_grad():
        bd = model.forward_features(points, frame_offsets, batch_size)
        n_loc, nc = model.dense_head.n_loc, model.num_class
        logits = bd["cls_preds"].reshape(-1, n_loc * nc)                    # rows = BEV locations, cols = conv_cls channels
        thr = float(np.log(model.cfg["score_thresh"] / (1 - model.cfg["score_thresh"])))
        # per output channel (anchor type x class): with random weights the channel MEANS differ far more than the
        # spatial variation, so anything coarser lets one channel win every arg-max
        mean, std = logits.mean(0), logits.std(0).clamp_min(1e-6)
        zs = ((logits - mean) / std).flatten()
        z = torch.quantile(zs[:: max(1, zs.numel() // 2000000)], 1.0 - target_fraction)
        w = model.dense_head.conv_cls.weight.view(n_loc * nc, -1)
        b = model.dense_head.conv_cls.bias.view(n_loc * nc)
        w /= std.view(-1, 1)                                                # new logit = (old - mean) / std - z + thr
        b.copy_((b - mean) / std - z + thr)
        if getattr(model.backbone_2d, "_plan", None) is not None:
            model.backbone_2d.build_inference_plan()
        if getattr(model.dense_head, "_plan", None) is not None:
            model.dense_head.build_inference_plan()
        if getattr(model, "_inference_args", None) is not None:
            model._inference_key = model._state_key()
    return model
