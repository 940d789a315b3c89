"""-m gpu parity of the anchor-head training path: crb3d target assignment and losses vs the REFERENCE's own
AxisAlignedTargetAssigner and AnchorHeadTemplate.get_loss (its unmodified classes, built by its SECONDNet from its second.yaml,
evaluated by torch on the same GPU)."""
import numpy as np
import pytest
import torch

import ref_env

pytestmark = pytest.mark.gpu


def _gt_batch(cuda, B=3, pad=40):
    from crb3d import synth
    gts = [synth.make_frame(10 + i, return_boxes=True)[1] for i in range(B)]
    gts[1] = gts[1][:0]                                             # a frame without any ground truth
    gt = np.zeros((B, pad, 8), np.float32)
    for i, g in enumerate(gts):
        gt[i, :len(g)] = g
    return torch.from_numpy(gt).to(cuda)


@pytest.fixture(scope="module")
def ref_head(cuda):
    if ref_env.install() is None:
        pytest.skip("no reference tree (neither /root/reference nor baseline/_ref)")
    reg = ref_env.register_model_families()
    cfg = ref_env.load_cfg("kitti_models/second.yaml")
    ds = ref_env.dataset_stub(cfg.DATA_CONFIG, cfg.CLASS_NAMES)
    torch.manual_seed(0)
    model = reg["SECONDNet"](model_cfg=cfg.MODEL, num_class=len(cfg.CLASS_NAMES), dataset=ds).cuda()
    return model.dense_head


def test_target_assignment_matches_reference(cuda, ref_head):
    from crb3d import second
    mine = second.SECONDNet().to_device(cuda).dense_head
    gt = _gt_batch(cuda)
    ref_t = ref_head.assign_targets(gt.clone())
    # anchors of this library in the reference's order. The reference builds its x / y centres with torch.arange(dtype=float32) on
    # the CPU, whose vectorised path adds lane * step in fp32 to a per-vector base (ATen RangeFactories: the values depend on the
    # host's SIMD width); this library evaluates start + i * step in float64 and rounds once. They agree to one ulp - the parity
    # checks below feed the reference's own anchors to both sides.
    ref_anchors = torch.cat(ref_head.anchors, dim=-3).view(-1, 7).contiguous()
    assert torch.allclose(mine.anchors_device(cuda), ref_anchors, rtol=2e-7, atol=1e-6)
    assert torch.equal(mine.anchors_device(cuda)[:, 2:], ref_anchors[:, 2:])
    labels_own_anchors = mine.assign_targets(gt)["box_cls_labels"]
    mine._anchors_dev = ref_anchors
    my_t = mine.assign_targets(gt)
    assert float((labels_own_anchors != my_t["box_cls_labels"]).float().mean()) < 1e-5     # one-ulp anchors move (almost) no label
    lab_r, lab_m = ref_t["box_cls_labels"].int(), my_t["box_cls_labels"]
    assert lab_m.shape == lab_r.shape == (3, mine.num_anchors)
    assert torch.equal(lab_m, lab_r)                                 # index-exact: -1 / 0 / class for all 3 x 211 200 anchors
    assert int((lab_m > 0).sum()) > 50 and int((lab_m == -1).sum()) > 0 and int((lab_m[1] != 0).sum()) == 0
    assert torch.equal(my_t["reg_weights"], ref_t["reg_weights"])
    # same fp32 expressions (division, log, sqrt): at most one ulp apart
    assert torch.allclose(my_t["box_reg_targets"], ref_t["box_reg_targets"], rtol=5e-7, atol=1e-7)
    assert torch.equal(my_t["num_pos"].long(), (lab_r > 0).sum(1))


def test_anchor_head_loss_and_gradients_match_reference(cuda, ref_head):
    from crb3d import second, train_ops
    mine = second.SECONDNet().to_device(cuda).dense_head
    mine._anchors_dev = torch.cat(ref_head.anchors, dim=-3).view(-1, 7).contiguous()       # see the note in the test above
    gt = _gt_batch(cuda)
    t = mine.assign_targets(gt)
    B, A = t["box_cls_labels"].shape
    g = torch.Generator(device=cuda).manual_seed(1)
    cls = (torch.randn(B, A, 3, device=cuda, generator=g) * 2 - 2).requires_grad_()
    # predictions near the targets for half of the anchors so that both smooth-L1 branches are exercised
    box = (t["box_reg_targets"] + torch.randn(B, A, 7, device=cuda, generator=g) * 0.2).detach().requires_grad_()
    dirp = torch.randn(B, A, 2, device=cuda, generator=g).requires_grad_()
    H, W = 200, 176                                                   # the reference views (B, H, W, anchors_per_location * C)
    ref_head.forward_ret_dict = {"cls_preds": cls.view(B, H, W, -1), "box_preds": box.view(B, H, W, -1),
                                 "dir_cls_preds": dirp.view(B, H, W, -1), "box_cls_labels": t["box_cls_labels"].clone(),
                                 "box_reg_targets": t["box_reg_targets"]}
    ref_loss, tb = ref_head.get_loss()
    ref_g = torch.autograd.grad(ref_loss, (cls, box, dirp))
    cls2, box2, dir2 = (x.detach().clone().requires_grad_() for x in (cls, box, dirp))
    cfg = dict(mine.cfg["loss"], dir_offset=mine.cfg["dir_offset"])
    losses = train_ops.anchor_head_loss(cls2, box2, dir2, t["box_cls_labels"], t["box_reg_targets"], mine.anchors_device(cuda), cfg)
    my_g = torch.autograd.grad(losses.sum(), (cls2, box2, dir2))
    assert abs(float(losses[0]) - tb["rpn_loss_cls"]) <= 1e-5 * abs(tb["rpn_loss_cls"]) + 1e-7
    assert abs(float(losses[1]) - tb["rpn_loss_loc"]) <= 1e-5 * abs(tb["rpn_loss_loc"]) + 1e-7
    assert abs(float(losses[2]) - tb["rpn_loss_dir"]) <= 1e-5 * abs(tb["rpn_loss_dir"]) + 1e-7
    assert abs(float(losses.sum()) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    for a, b, name in zip(my_g, ref_g, ("cls", "box", "dir")):
        scale = float(b.abs().max())
        assert scale > 0
        assert float((a - b).abs().max()) <= 1e-5 * scale + 1e-9, name
    # module-level path: forward in training mode assigns targets, get_loss gives the same total
    mine.train()
    mine.forward_ret_dict = {"cls_preds": cls2, "box_preds": box2, "dir_cls_preds": dir2, **t}
    total, tb2 = mine.get_loss()
    assert abs(float(total) - float(ref_loss)) <= 1e-5 * abs(float(ref_loss))
    assert set(tb2) == {"rpn_loss_cls", "rpn_loss_loc", "rpn_loss_dir", "rpn_loss"}
    # NaN regression targets are ignored (loss_utils.py:123)
    rt = t["box_reg_targets"].clone()
    pos = (t["box_cls_labels"] > 0).nonzero()[:3]
    rt[pos[:, 0], pos[:, 1], 2] = float("nan")
    l_nan = train_ops.anchor_head_loss(cls2, box2, dir2, t["box_cls_labels"], rt, mine.anchors_device(cuda), cfg)
    assert torch.isfinite(l_nan).all() and float(l_nan[1]) <= float(losses[1]) + 1e-6


def test_train_step_with_targets_runs_and_decreases_loss(cuda):
    """SECOND forward + backward with the device-side target assignment and losses: a few SGD steps on one batch lower rpn_loss."""
    from crb3d import second, synth
    torch.manual_seed(0)
    model = second.SECONDNet().to_device(cuda).train()
    pts, offs, _ = synth.make_batch([20, 21])
    pts, offs = torch.from_numpy(pts).to(cuda), torch.from_numpy(offs).to(cuda)
    gts = [synth.make_frame(i, return_boxes=True)[1] for i in (20, 21)]
    gt = np.zeros((2, max(len(g) for g in gts), 8), np.float32)
    for i, g in enumerate(gts):
        gt[i, :len(g)] = g
    gt = torch.from_numpy(gt).to(cuda)
    opt = torch.optim.SGD(model.parameters(), lr=2e-3, momentum=0.9)
    hist = []
    for _ in range(6):
        opt.zero_grad(set_to_none=True)
        model.forward_features(pts, offs, 2, gt_boxes=gt)
        loss, tb = model.dense_head.get_loss()
        loss.backward()
        opt.step()
        hist.append(float(loss))
    assert all(np.isfinite(hist)) and hist[-1] < hist[0], hist


def test_target_assignment_without_ground_truth(cuda):
    """No ground truth at all (M = 0) and only padding rows: every anchor is background, no regression target, no positive."""
    from crb3d import second
    head = second.SECONDNet().to_device(cuda).dense_head
    for gt in (torch.zeros((2, 0, 8), device=cuda), torch.zeros((2, 5, 8), device=cuda)):
        t = head.assign_targets(gt)
        assert int(t["box_cls_labels"].abs().sum()) == 0 and float(t["box_reg_targets"].abs().sum()) == 0.0
        assert int(t["num_pos"].sum()) == 0 and float(t["reg_weights"].sum()) == 0.0
