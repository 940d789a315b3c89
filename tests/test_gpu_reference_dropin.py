"""-m gpu: the REFERENCE's own, unmodified Python files running on top of the crb3d drop-in.

tests/ref_env.py imports them from /root/reference or from the verbatim staging baseline/_ref (git-ignored, travels with
gpurun): op wrappers, spconv_backbone.VoxelBackBone8x, MeanVFE, HeightCompression, BaseBEVBackbone, AnchorHeadSingle, the
SECONDNet / PVRCNN detector classes built from the reference's own YAML configs - `import spconv.pytorch`, `from . import
iou3d_nms_cuda` / `roiaware_pool3d_cuda` / `pointnet2_stack_cuda` all resolve to this library.
"""
import numpy as np
import pytest
import torch

import ref_env
from util import cu, rand_boxes

pytestmark = pytest.mark.gpu
needs_ref = pytest.mark.skipif(ref_env.reference_root() is None, reason="no reference tree (/root/reference or baseline/_ref)")


def _batch(cuda, n_frames, cfg_data, with_voxels=True, max_voxels=40000):
    """pcdet-style batch_dict of synthetic frames: points (N, 1+C) with the batch index in column 0 (dataset.py:180-186),
    hard voxels from this library's voxelizer (the reference's DataProcessor calls spconv's through the same shim)."""
    from crb3d import ops, synth
    frames = [synth.make_frame(i) for i in range(n_frames)]
    offs = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    pts = torch.from_numpy(np.concatenate(frames)).to(cuda)
    offs_t = torch.from_numpy(offs).to(cuda)
    bidx = torch.repeat_interleave(torch.arange(n_frames, device=cuda), torch.from_numpy(np.diff(offs)).to(cuda)).float()
    bd = dict(batch_size=n_frames, points=torch.cat([bidx[:, None], pts], 1).contiguous(),
              frame_id=np.arange(n_frames), gt_boxes=torch.zeros((n_frames, 1, 8), device=cuda))
    gt = [synth.make_boxes(np.random.default_rng(1000 + i), synth.KITTI) for i in range(n_frames)]
    g = max(len(x) for x in gt)
    gtb = np.zeros((n_frames, g, 8), np.float32)
    for i, x in enumerate(gt):
        gtb[i, :len(x)] = x
    bd["gt_boxes"] = torch.from_numpy(gtb).to(cuda)
    if with_voxels:
        vox = ops.voxelize(pts, offs_t, n_frames, cfg_data.POINT_CLOUD_RANGE, [0.05, 0.05, 0.1], 5, max_voxels, want_voxels=True)
        bd.update(voxels=vox["voxels"], voxel_num_points=vox["num_points"], voxel_coords=vox["coords"].float())
    return bd, frames, pts, offs_t


@needs_ref
def test_reference_op_wrappers_run_unmodified_on_dropin(cuda):
    """pcdet/ops/*/..._utils.py of the REFERENCE imported as they are."""
    iou_utils = ref_env.ref("pcdet.ops.iou3d_nms.iou3d_nms_utils")
    roi_utils = ref_env.ref("pcdet.ops.roiaware_pool3d.roiaware_pool3d_utils")
    pn_utils = ref_env.ref("pcdet.ops.pointnet2.pointnet2_stack.pointnet2_utils")
    assert iou_utils.iou3d_nms_cuda.__name__ == "pcdet_ops.iou3d_nms_cuda"
    from crb3d import box_ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(3)
    a, b = cu(rand_boxes(rng, 80, 10, True), cuda), cu(rand_boxes(rng, 60, 10, True), cuda)
    assert torch.equal(iou_utils.boxes_iou3d_gpu(a, b), box_ops.boxes_iou3d_gpu(a, b))
    scores = cu(rng.permutation(80).astype(np.float32), cuda)
    k_ref, _ = iou_utils.nms_gpu(a, scores, 0.25)
    k_mine, _ = box_ops.nms_gpu(a, scores, 0.25)
    assert torch.equal(k_ref, k_mine)
    pts = cu(rng.uniform(-12, 12, (2, 500, 3)).astype(np.float32), cuda)
    bx = torch.stack([a[:30], b[:30]])
    assert torch.equal(roi_utils.points_in_boxes_gpu(pts, bx), box_ops.points_in_boxes_gpu(pts, bx))
    # the reference's BallQuery / FPS autograd functions over pointnet2_stack_cuda == the oracle's restatement
    xyz = rng.uniform(-6, 6, (3000, 3)).astype(np.float32)
    new_xyz = xyz[:200] + np.float32(0.03)
    cnt, ncnt = np.array([3000], np.int32), np.array([200], np.int32)
    idx, empty = pn_utils.ball_query(0.8, 16, cu(xyz, cuda), cu(cnt, cuda), cu(new_xyz, cuda), cu(ncnt, cuda))
    idx_o = op.ball_query(0.8, 16, xyz, cnt, new_xyz, ncnt)
    idx_o[idx_o[:, 0] == -1] = 0
    assert np.array_equal(idx.cpu().numpy(), idx_o)
    fps = pn_utils.farthest_point_sample(cu(xyz[None], cuda), 64)
    assert np.array_equal(fps[0].cpu().numpy(), op.farthest_point_sampling(xyz, 64)[0])


@needs_ref
def test_reference_second_modules_over_dropin_equal_crb3d(cuda):
    """The reference's VoxelBackBone8x + HeightCompression + BaseBEVBackbone + AnchorHeadSingle (built by ITS SECONDNet class
    from ITS second.yaml) over the spconv shim vs crb3d.second.SECONDNet with the same state_dict: the sparse backbone has
    identical rows and activations within 1e-5 (same conv kernels; BatchNorm fused vs separate), the dense stack agrees to fp32 round-off (NCHW vs channels-last cuDNN algorithms),
    the decoded boxes agree with the fused head kernels, and the reference checkpoint keys load into crb3d.second unchanged."""
    from crb3d import head_ops, ops, second
    reg = ref_env.register_model_families()
    cfg = ref_env.load_cfg("kitti_models/second.yaml")
    ds = ref_env.dataset_stub(cfg.DATA_CONFIG, cfg.CLASS_NAMES)
    torch.manual_seed(0)
    ref_model = reg["SECONDNet"](model_cfg=cfg.MODEL, num_class=len(cfg.CLASS_NAMES), dataset=ds).cuda().eval()
    mine = second.SECONDNet().eval().to_device(cuda)
    missing, unexpected = mine.load_state_dict(ref_model.state_dict(), strict=False)
    assert not [k for k in missing if "num_batches_tracked" not in k], missing
    assert not [k for k in unexpected if "global_step" not in k and "num_batches_tracked" not in k], unexpected
    bd, frames, pts, offs_t = _batch(cuda, 2, cfg.DATA_CONFIG)
    old_tf32, old_sp = torch.backends.cudnn.allow_tf32, ops.SPCONV_TF32
    torch.backends.cudnn.allow_tf32 = False
    ops.SPCONV_TF32 = False
    try:
        with torch.no_grad():
            rd = dict(bd)
            for m in ref_model.module_list:      # SECONDNet.forward without post_processing (it needs RoI-head keys, SURVEY 8a note G)
                rd = m(rd)
            md = mine.forward_features(pts, offs_t, 2)
    finally:
        torch.backends.cudnn.allow_tf32, ops.SPCONV_TF32 = old_tf32, old_sp
    enc_r, enc_m = rd["encoded_spconv_tensor"], md["encoded_spconv_tensor"]
    def close(a, b, rel=1e-5):      # same conv kernels; the reference applies BatchNorm1d + ReLU as separate torch ops,
        return float((a - b).abs().max()) <= rel * float(a.abs().max())   # crb3d fuses them into the conv epilogue (one fma)
    assert torch.equal(enc_r.indices, enc_m.indices) and close(enc_r.features, enc_m.features)
    for k in ("x_conv1", "x_conv2", "x_conv3", "x_conv4"):
        assert torch.equal(rd["multi_scale_3d_features"][k].indices, md["multi_scale_3d_features"][k].indices)
        assert close(rd["multi_scale_3d_features"][k].features, md["multi_scale_3d_features"][k].features)
    assert close(rd["spatial_features"], md["spatial_features"].contiguous())
    s = float(rd["spatial_features_2d"].abs().max())
    assert float((rd["spatial_features_2d"] - md["spatial_features_2d"]).abs().max()) <= 1e-4 * s
    # reference decode (anchor_head_template.generate_predicted_boxes) vs the fused decode kernel on the same head outputs
    A = mine.dense_head.num_anchors
    idx = torch.arange(A, device=cuda).view(1, A).repeat(2, 1)
    boxes = head_ops.anchor_decode_select(md["box_preds"], md["dir_cls_preds"], idx, mine.dense_head.spec, A)
    assert float((boxes - rd["batch_box_preds"]).abs().max()) <= 2e-3     # sin/cos/exp of values that differ by 1e-4 relative
    assert float((md["cls_preds"] - rd["batch_cls_preds"]).abs().max()) <= 1e-4 * float(rd["batch_cls_preds"].abs().max()) + 1e-5


@needs_ref
def test_reference_pvrcnn_detector_forward_over_dropin(cuda):
    """The reference's PVRCNN class built from ITS pv_rcnn_active_crb.yaml runs its whole eval forward - MeanVFE, spconv
    backbone, BEV backbone, anchor head, VoxelSetAbstraction (FPS + 5 StackSAModuleMSG), PointHeadSimple, PVRCNNHead
    (proposal NMS, RoI-grid pooling, 5 MC-dropout rounds) and post_processing (NMS, points-in-boxes density) - with every
    compiled op answered by this library; its records are what CRBSampling.query consumes (crb_sampling.py:72-103)."""
    reg = ref_env.register_model_families()
    cfg = ref_env.load_cfg("active-kitti_models/pv_rcnn_active_crb.yaml")
    ref_env.set_global_cfg(cfg)
    ds = ref_env.dataset_stub(cfg.DATA_CONFIG, cfg.CLASS_NAMES)
    torch.manual_seed(0)
    model = reg["PVRCNN"](model_cfg=cfg.MODEL, num_class=len(cfg.CLASS_NAMES), dataset=ds).cuda().eval()
    bd, frames, pts, offs_t = _batch(cuda, 2, cfg.DATA_CONFIG)
    with torch.no_grad():
        pred_dicts, _ = model(bd)
    assert len(pred_dicts) == 2
    for p in pred_dicts:
        n = p["pred_boxes"].shape[0]
        assert p["pred_labels"].shape[0] == n and p["pred_box_unique_density"].shape[0] == n
        assert torch.isfinite(p["pred_boxes"]).all() and torch.isfinite(p["pred_box_unique_density"]).all()
        assert p["batch_rcnn_cls"].shape[-1] == 1 and p["batch_rcnn_reg"].shape[-1] == 7
