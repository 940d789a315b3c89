"""-m gpu: the host-side mirrors of the reference's op wrappers (crb3d.pointnet2_modules, crb3d.box_ops) - forward values
vs the oracle, gradients vs torch autograd over an index-based restatement, and the reference's own wrapper files
running unmodified on top of the drop-in modules when /root/reference is present."""
import os

import numpy as np
import pytest
import torch

from util import cu, rand_boxes

pytestmark = pytest.mark.gpu


def test_query_and_group_and_sa_module(cuda):
    from crb3d import pointnet2_modules as pm
    from oracle import pointnet2 as op
    rng = np.random.default_rng(0)
    xyz_cnt, new_cnt = np.array([3000, 2500], np.int32), np.array([300, 200], np.int32)
    xyz = rng.uniform(-6, 6, (5500, 3)).astype(np.float32)
    new_xyz = np.concatenate([xyz[:300] + np.float32(0.05), xyz[3000:3200] + np.float32(0.05)])
    new_xyz[7] = 100.0                                                    # an empty ball
    feat = rng.normal(size=(5500, 16)).astype(np.float32)
    f = cu(feat, cuda).requires_grad_(True)
    qg = pm.QueryAndGroup(0.8, 16, use_xyz=True)
    out, idx = qg(cu(xyz, cuda), cu(xyz_cnt, cuda), cu(new_xyz, cuda), cu(new_cnt, cuda), f)
    idx_o = op.ball_query(0.8, 16, xyz, xyz_cnt, new_xyz, new_cnt)
    empty = idx_o[:, 0] == -1
    idx_o[empty] = 0
    assert np.array_equal(idx.cpu().numpy(), idx_o) and empty[7]
    g_xyz = op.group_points(xyz, xyz_cnt, idx_o, new_cnt) - new_xyz[:, :, None]
    g_f = op.group_points(feat, xyz_cnt, idx_o, new_cnt)
    g_xyz[empty] = 0
    g_f[empty] = 0
    assert np.allclose(out.detach().cpu().numpy(), np.concatenate([g_xyz, g_f], 1), atol=1e-6)
    w = torch.randn(out.shape, device=cuda)
    (out * w).sum().backward()
    # gradient of the gather == scatter-add of the upstream gradient
    starts = np.concatenate([np.zeros(300, np.int64), np.full(200, 3000, np.int64)])
    gi = torch.from_numpy(idx_o.astype(np.int64) + starts[:, None]).to(cuda)
    ref = torch.zeros_like(f)
    wf = w[:, 3:, :].clone()
    wf[torch.from_numpy(empty).to(cuda)] = 0
    ref.index_add_(0, gi.reshape(-1), wf.permute(0, 2, 1).reshape(-1, 16))
    assert torch.allclose(f.grad, ref, rtol=1e-4, atol=1e-4)
    sa = pm.StackSAModuleMSG([0.4, 0.8], [16, 16], [[16, 16, 16], [16, 32, 32]]).to(cuda).eval()
    _, pooled = sa(cu(xyz, cuda), cu(xyz_cnt, cuda), cu(new_xyz, cuda), cu(new_cnt, cuda), cu(feat, cuda))
    assert pooled.shape == (500, 48) and torch.isfinite(pooled).all()


def test_fps_three_nn_wrappers(cuda):
    from crb3d import pointnet2_modules as pm
    from oracle import pointnet2 as op
    rng = np.random.default_rng(1)
    pts = rng.uniform(-10, 10, (2, 4000, 3)).astype(np.float32)
    idx = pm.farthest_point_sample(cu(pts, cuda), 256)
    for b in range(2):
        assert np.array_equal(idx[b].cpu().numpy(), op.farthest_point_sampling(pts[b], 256)[0])
    cnt = torch.tensor([4000, 4000], dtype=torch.int32, device=cuda)
    sidx = pm.stack_farthest_point_sample(cu(pts.reshape(-1, 3), cuda), cnt, 128)
    assert np.array_equal(sidx[:128].cpu().numpy(), op.farthest_point_sampling(pts[0], 128, block=1024)[0])
    unknown, known = rng.uniform(-5, 5, (900, 3)).astype(np.float32), rng.uniform(-5, 5, (200, 3)).astype(np.float32)
    uc, kc = torch.tensor([500, 400], dtype=torch.int32, device=cuda), torch.tensor([120, 80], dtype=torch.int32, device=cuda)
    dist, nidx = pm.three_nn(cu(unknown, cuda), uc, cu(known, cuda), kc)
    d_o, i_o = op.three_nn(unknown, [500, 400], known, [120, 80])
    assert np.array_equal(nidx.cpu().numpy(), i_o) and np.allclose(dist.cpu().numpy(), np.sqrt(d_o), rtol=1e-6)
    feats = cu(rng.normal(size=(200, 8)).astype(np.float32), cuda).requires_grad_(True)
    wgt = torch.softmax(torch.randn(900, 3, device=cuda), -1)
    out = pm.three_interpolate(feats, nidx, wgt)
    ref = (feats[nidx.long()] * wgt.unsqueeze(-1)).sum(1)
    assert torch.allclose(out, ref, rtol=1e-5, atol=1e-6)
    out.sum().backward()
    g = torch.autograd.grad(ref.sum(), feats)[0]
    assert torch.allclose(feats.grad, g, rtol=1e-4, atol=1e-4)


def test_box_ops_wrappers(cuda):
    from crb3d import box_ops
    from oracle import boxes as ob
    rng = np.random.default_rng(2)
    a, b = rand_boxes(rng, 200, 12, True), rand_boxes(rng, 150, 12, True)
    iou3d = box_ops.boxes_iou3d_gpu(cu(a, cuda), cu(b, cuda)).cpu().numpy()
    assert np.abs(iou3d - ob.boxes_iou3d(a, b)).max() < 1e-4
    scores = rng.permutation(200).astype(np.float32)                       # untied scores
    keep, _ = box_ops.nms_gpu(cu(a, cuda), cu(scores, cuda), 0.3)
    order = np.argsort(-scores)
    keep_o, iou = ob.nms_sorted(a[order], 0.3, return_iou=True)
    if np.abs(iou[np.triu_indices(200, 1)] - np.float32(0.3)).min() > 1e-5:
        assert np.array_equal(keep.cpu().numpy(), order[keep_o])
    sel, sc = box_ops.class_agnostic_nms(cu(scores / 200, cuda), cu(a, cuda),
                                         dict(NMS_TYPE="nms_gpu", NMS_THRESH=0.3, NMS_PRE_MAXSIZE=150, NMS_POST_MAXSIZE=20), 0.2)
    assert len(sel) <= 20 and float(sc.min()) >= 0.2
    pool = box_ops.RoIAwarePool3d(out_size=4, max_pts_each_voxel=16)
    pts = cu(rng.uniform(-12, 12, (3000, 3)).astype(np.float32), cuda)
    feat = cu(rng.normal(size=(3000, 8)).astype(np.float32), cuda).requires_grad_(True)
    out = pool(cu(a[:10], cuda), pts, feat, pool_method="max")
    p_o, _, _ = ob.roiaware_pool3d(a[:10], pts.cpu().numpy(), feat.detach().cpu().numpy(), 4, 16, "max")
    assert (np.abs(out.detach().cpu().numpy() - p_o) > 1e-6).mean() < 1e-3
    out.sum().backward()
    assert feat.grad.abs().sum() > 0
