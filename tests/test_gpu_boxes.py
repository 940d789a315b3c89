"""-m gpu parity: rotated IoU / NMS / points-in-boxes / RoI-aware pool - ours vs the C oracle and, bit for bit, vs the
reference's own CUDA kernels (oracle/_ref) running on the same device."""
import ctypes

import numpy as np
import pytest
import torch

from util import P, cu, rand_boxes, ref_kernels

pytestmark = pytest.mark.gpu


def _ref_or_skip():
    ref = ref_kernels()
    if ref is None:
        pytest.skip("oracle/_ref/libpcdet_ref_kernels.so not built (needs /root/reference at build time)")
    return ref


@pytest.mark.parametrize("cluster", [False, True])
def test_pairwise_overlap_iou_vs_oracle(cuda, cluster):
    from crb3d import ops
    from oracle import boxes as ob
    rng = np.random.default_rng(3 + cluster)
    a, b = rand_boxes(rng, 300, 20, cluster), rand_boxes(rng, 200, 20, cluster)
    ov = ops.boxes_overlap_bev(cu(a, cuda), cu(b, cuda)).cpu().numpy()
    iou = ops.boxes_iou_bev(cu(a, cuda), cu(b, cuda)).cpu().numpy()
    # CPU libm vs CUDA sincos/atan2 differ by ulps -> tolerance 1e-4 absolute on areas of O(10), 1e-5 on IoU
    assert np.abs(ov - ob.boxes_overlap_bev(a, b)).max() < 1e-3
    assert np.abs(iou - ob.boxes_iou_bev(a, b)).max() < 1e-4
    assert (ov > 0).mean() > (0.02 if cluster else 0.001)


def test_pairwise_bitwise_vs_reference_kernels(cuda):
    from crb3d import ops
    ref = _ref_or_skip()
    rng = np.random.default_rng(17)
    mism = 0
    for cluster in (False, True):
        a, b = cu(rand_boxes(rng, 1500, 25, cluster), cuda), cu(rand_boxes(rng, 1300, 25, cluster), cuda)
        mine_ov, mine_iou = ops.boxes_overlap_bev(a, b), ops.boxes_iou_bev(a, b)
        r_ov, r_iou = torch.zeros_like(mine_ov), torch.zeros_like(mine_iou)
        torch.cuda.synchronize()
        ref.ref_boxes_overlap(a.shape[0], P(a), b.shape[0], P(b), P(r_ov))
        ref.ref_boxes_iou_bev(a.shape[0], P(a), b.shape[0], P(b), P(r_iou))
        assert ref.ref_sync() == 0
        mism += int((mine_ov != r_ov).sum()) + int((mine_iou != r_iou).sum())
    assert mism == 0, "%d of ~7.8M pair values differ from the reference kernels bitwise" % mism


@pytest.mark.parametrize("n,thr,rotated", [(4096, 0.01, True), (1024, 0.7, True), (9000, 0.8, True), (777, 0.1, True),
                                           (2000, 0.5, False), (1, 0.5, True), (64, 0.3, True), (65, 0.3, True),
                                           (15000, 0.5, True), (513, -1.0, True)])
def test_nms_vs_reference_kernels(cuda, n, thr, rotated):
    """keep list == the reference's mask kernel + the host greedy loop of iou3d_nms.cpp:116-132 (restated in numpy)."""
    from crb3d import ops
    ref = _ref_or_skip()
    rng = np.random.default_rng(n)
    boxes = cu(rand_boxes(rng, n, 40, cluster=True), cuda)
    keep, num = ops.nms_sorted(boxes, thr, rotated=rotated)
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=cuda)
    torch.cuda.synchronize()
    (ref.ref_nms_mask if rotated else ref.ref_nms_normal_mask)(P(boxes), P(mask), n, ctypes.c_float(thr))
    assert ref.ref_sync() == 0
    m = mask.cpu().numpy().view(np.uint64)
    remv = np.zeros(cb, np.uint64)
    ref_keep = []
    for i in range(n):
        if not (int(remv[i // 64]) >> (i % 64)) & 1:
            ref_keep.append(i)
            remv[i // 64:] |= m[i, i // 64:]
    got = keep[: int(num.item())].cpu().numpy()
    assert np.array_equal(got, np.asarray(ref_keep, np.int64))
    # upper-triangular tiles of our mask equal the reference's, bit for bit
    mine = ops.nms_mask(boxes, thr, rotated).cpu().numpy().view(np.uint64)
    for r in range(cb):
        assert np.array_equal(mine[r * 64:(r + 1) * 64, r:], m[r * 64:(r + 1) * 64, r:])
    if n > 500:
        k500, n500 = ops.nms_sorted(boxes, thr, rotated=rotated, max_keep=500)
        assert np.array_equal(k500[: int(n500.item())].cpu().numpy(), np.asarray(ref_keep[:500], np.int64))


def test_nms_vs_oracle_with_margin(cuda):
    from crb3d import ops
    from oracle import boxes as ob
    rng = np.random.default_rng(23)
    b = rand_boxes(rng, 600, 15, cluster=True)
    _, iou = ob.nms_sorted(b, 0.3, return_iou=True)
    vals = iou[np.triu_indices(600, 1)]
    # pick a threshold no pair sits on (CPU libm vs CUDA sincos/atan2 differ by ulps, so a margin is required)
    thr = next(t for t in np.arange(0.30, 0.40, 0.003) if np.abs(vals - np.float32(t)).min() > 2e-5)
    keep_o = ob.nms_sorted(b, thr)
    keep, num = ops.nms_sorted(cu(b, cuda), thr)
    assert np.array_equal(keep[: int(num.item())].cpu().numpy(), keep_o)


def _pib_inputs(rng, B, T, M):
    boxes = np.stack([rand_boxes(rng, T, 20) for _ in range(B)])
    pts = rng.uniform([-22, -22, -2.5], [22, 22, 1.5], (B, M, 3)).astype(np.float32)
    # adversarial: points on / next to box faces in the box frame
    for b in range(B):
        for t in range(min(T, 40)):
            bx = boxes[b, t]
            c, s = np.cos(bx[6]), np.sin(bx[6])
            for j, (lx, ly, lz) in enumerate([(bx[3] / 2, 0, 0), (0, bx[4] / 2, 0), (0, 0, bx[5] / 2), (-bx[3] / 2, -bx[4] / 2, -bx[5] / 2)]):
                k = (t * 4 + j) % M
                pts[b, k] = [bx[0] + lx * c - ly * s, bx[1] + lx * s + ly * c, bx[2] + lz]
    return boxes, pts


def test_points_in_boxes_vs_oracle_and_reference(cuda):
    from crb3d import ops
    from oracle import boxes as ob
    rng = np.random.default_rng(31)
    B, T, M = 3, 150, 20000
    boxes, pts = _pib_inputs(rng, B, T, M)
    out = ops.points_in_boxes(cu(boxes, cuda), cu(pts, cuda)).cpu().numpy()
    assert (out >= 0).mean() > 0.05
    ref = ref_kernels()
    if ref is not None:
        r = torch.full((B, M), -1, dtype=torch.int32, device=cuda)
        bt, pt = cu(boxes, cuda), cu(pts, cuda)
        torch.cuda.synchronize()
        ref.ref_points_in_boxes(B, T, M, P(bt), P(pt), P(r))
        assert ref.ref_sync() == 0
        assert np.array_equal(out, r.cpu().numpy())                 # bit-exact vs the reference kernel
    # CPU oracle: sinf/cosf of libm vs CUDA may differ in the last ulp -> allow only on-the-face points to differ
    o = np.stack([ob.points_in_boxes(boxes[b], pts[b]) for b in range(B)])
    assert (out != o).mean() < 2e-4


def test_points_in_boxes_stack_density(cuda):
    from crb3d import ops
    from oracle import boxes as ob
    rng = np.random.default_rng(37)
    frames = [rng.uniform([-22, -22, -2.5, 0], [22, 22, 1.5, 1], (n, 4)).astype(np.float32) for n in (15000, 0, 18000)]
    fboxes = [rand_boxes(rng, t, 20) for t in (60, 10, 0)]
    pt_off = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    box_off = np.cumsum([0] + [len(b) for b in fboxes]).astype(np.int32)
    idx, counts, dens = ops.points_in_boxes_stack(cu(np.concatenate(frames), cuda), cu(pt_off, cuda),
                                                  cu(np.concatenate(fboxes), cuda), cu(box_off, cuda), 18000)
    for b in range(3):
        d_o, c_o, i_o = ob.box_density(fboxes[b], frames[b])
        assert np.array_equal(idx[pt_off[b]:pt_off[b + 1]].cpu().numpy(), i_o)
        assert np.array_equal(counts[box_off[b]:box_off[b + 1]].cpu().numpy(), c_o)
        assert np.array_equal(dens[box_off[b]:box_off[b + 1]].cpu().numpy(), d_o)   # count/(dx*dy*dz), same fp32 ops


@pytest.mark.parametrize("method", ["max", "avg"])
def test_roiaware_pool3d(cuda, method):
    from crb3d import ops
    from oracle import boxes as ob
    rng = np.random.default_rng(41)
    n_boxes, n_pts, C, out, mp = 24, 6000, 16, 6, 16
    rois = rand_boxes(rng, n_boxes, 10)
    pts = rng.uniform([-12, -12, -2.5], [12, 12, 1.5], (n_pts, 3)).astype(np.float32)
    feat = rng.normal(size=(n_pts, C)).astype(np.float32)
    pooled = torch.zeros((n_boxes, out, out, out, C), device=cuda)
    argmax = torch.zeros((n_boxes, out, out, out, C), dtype=torch.int32, device=cuda)
    pidx = torch.zeros((n_boxes, out, out, out, mp), dtype=torch.int32, device=cuda)
    m = {"max": 0, "avg": 1}[method]
    ops.roiaware_pool3d_forward(cu(rois, cuda), cu(pts, cuda), cu(feat, cuda), argmax, pidx, pooled, m)
    p_o, a_o, i_o = ob.roiaware_pool3d(rois, pts, feat, out, mp, method)
    assert (i_o[..., 0] > 0).mean() > 0.05
    ref = ref_kernels()
    if ref is not None:
        rp, ra, ri = torch.zeros_like(pooled), torch.zeros_like(argmax), torch.zeros_like(pidx)
        rt, pt, ft = cu(rois, cuda), cu(pts, cuda), cu(feat, cuda)
        torch.cuda.synchronize()
        ref.ref_roiaware_pool3d(n_boxes, n_pts, C, mp, out, out, out, P(rt), P(pt), P(ft), P(ra), P(ri), P(rp), m)
        assert ref.ref_sync() == 0
        assert torch.equal(pidx, ri) and torch.equal(pooled, rp)
        if method == "max":
            assert torch.equal(argmax, ra)
        g = cu(rng.normal(size=tuple(pooled.shape)).astype(np.float32), cuda)
        gi, rgi = torch.zeros((n_pts, C), device=cuda), torch.zeros((n_pts, C), device=cuda)
        ops.roiaware_pool3d_backward(pidx, argmax, g, gi, m)
        torch.cuda.synchronize()
        ref.ref_roiaware_pool3d_backward(n_boxes, out, out, out, C, mp, P(ri), P(ra), P(g), P(rgi), m)
        assert ref.ref_sync() == 0
        assert torch.allclose(gi, rgi, rtol=1e-5, atol=1e-6)      # float atomics: order differs run to run
    same = (pidx.cpu().numpy() == i_o).mean()
    assert same > 0.9995                                            # libm-vs-CUDA sincos ulps on face points only
    if same == 1.0:
        assert np.allclose(pooled.cpu().numpy(), p_o, rtol=1e-6, atol=1e-6)
