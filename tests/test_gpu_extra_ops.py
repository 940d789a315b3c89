"""-m gpu parity of the remaining op families: voxel query, the pointnet2_batch layout and
roipoint_pool3d - ours vs the C oracle and vs the reference's own CUDA kernels (oracle/_ref/libpcdet_ref_kernels_batch.so),
plus the reference's unmodified Python wrappers running over the drop-in."""
import ctypes

import numpy as np
import pytest
import torch

import ref_env
from util import P, cu, rand_boxes, ref_batch_kernels

pytestmark = pytest.mark.gpu


def _cloud(rng, n, spread=8.0):
    p = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    p[:, 2] = rng.uniform(-2, 1, n)
    return p


@pytest.mark.parametrize("radius,nsample,n,m", [(0.4, 16, 5000, 300), (0.8, 32, 4096, 1024), (2.4, 64, 1500, 37), (1.0, 16, 31, 5)])
def test_ball_query_batch(cuda, radius, nsample, n, m):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(n + m)
    B = 3
    xyz = np.stack([_cloud(rng, n) for _ in range(B)])
    new_xyz = np.stack([np.concatenate([xyz[b, :m // 2] + np.float32(0.03), _cloud(rng, m - m // 2, 30.0)]) for b in range(B)])
    # sources at distance == radius (within rounding) of query 0 of frame 0
    k = min(64, n - 1)
    ang = rng.uniform(0, 2 * np.pi, k)
    xyz[0, 1:1 + k] = new_xyz[0, 0] + np.float32(radius) * np.stack([np.cos(ang), np.sin(ang), np.zeros(k)], 1).astype(np.float32)
    idx = torch.zeros((B, m, nsample), dtype=torch.int32, device=cuda)
    a, c = cu(new_xyz, cuda), cu(xyz, cuda)
    ops.ball_query_batch(B, n, m, radius, nsample, a, c, idx)
    # oracle: the stacked restatement frame by frame; the batch kernel has no -1 marker (an empty ball keeps the zeros)
    o = np.stack([op.ball_query(radius, nsample, xyz[b], np.array([n], np.int32), new_xyz[b], np.array([m], np.int32)) for b in range(B)])
    assert (o[..., 0] == -1).any() and (o[..., 0] >= 0).any()
    o[o[..., 0] == -1] = 0
    assert np.array_equal(idx.cpu().numpy(), o)
    ref = ref_batch_kernels()
    if ref is not None:
        r = torch.zeros_like(idx)
        torch.cuda.synchronize()
        ref.refb_ball_query(B, n, m, ctypes.c_float(radius), nsample, P(a), P(c), P(r))
        assert ref.refb_sync() == 0
        assert torch.equal(idx, r)


@pytest.mark.parametrize("C,ns", [(1, 16), (19, 8), (64, 32)])
def test_group_and_gather_points_batch_and_grads(cuda, C, ns):
    from crb3d import ops
    rng = np.random.default_rng(C * 7 + ns)
    B, n, npts = 2, 2000, 300
    pts = rng.normal(size=(B, C, n)).astype(np.float32)
    idx = rng.integers(0, n, (B, npts, ns)).astype(np.int32)
    pt, it = cu(pts, cuda), cu(idx, cuda)
    out = torch.zeros((B, C, npts, ns), device=cuda)
    ops.group_points_batch(B, C, n, npts, ns, pt, it, out)
    expect = np.stack([pts[b][:, idx[b]] for b in range(B)])              # (B, C, npts, ns)
    assert np.array_equal(out.cpu().numpy(), expect)
    g = rng.normal(size=(B, C, npts, ns)).astype(np.float32)
    gp = torch.zeros((B, C, n), device=cuda)
    ops.group_points_grad_batch(B, C, n, npts, ns, cu(g, cuda), it, gp)
    eg = np.zeros((B, C, n), np.float64)
    for b in range(B):
        for c in range(C):
            np.add.at(eg[b, c], idx[b].ravel(), g[b, c].ravel().astype(np.float64))
    assert np.allclose(gp.cpu().numpy(), eg, rtol=1e-4, atol=1e-5)
    # gather_points = the nsample 1 case (sampling_gpu.cu:15-70)
    from pcdet_ops import pointnet2_batch_cuda as shim
    gi = cu(idx[:, :, 0], cuda)
    go = torch.zeros((B, C, npts), device=cuda)
    shim.gather_points_wrapper(B, C, n, npts, pt, gi, go)
    assert np.array_equal(go.cpu().numpy(), np.stack([pts[b][:, idx[b, :, 0]] for b in range(B)]))
    ref = ref_batch_kernels()
    if ref is not None:
        r = torch.zeros_like(out)
        rg = torch.zeros_like(go)
        torch.cuda.synchronize()
        ref.refb_group_points(B, C, n, npts, ns, P(pt), P(it), P(r))
        ref.refb_gather_points(B, C, n, npts, P(pt), P(gi), P(rg))
        assert ref.refb_sync() == 0
        assert torch.equal(out, r) and torch.equal(go, rg)


@pytest.mark.parametrize("n,m,C", [(3000, 500, 32), (257, 3, 5), (1000, 2048, 128)])
def test_three_nn_and_interpolate_batch(cuda, n, m, C):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(n + m + C)
    B = 2
    unknown = np.stack([_cloud(rng, n) for _ in range(B)])
    known = np.stack([_cloud(rng, m) for _ in range(B)])
    if m > 20:
        known[0, 10:15] = known[0, 5:10]          # exact ties: the lowest index must win
        unknown[0, :5] = known[0, 5:10]
    u, k = cu(unknown, cuda), cu(known, cuda)
    d2 = torch.zeros((B, n, 3), device=cuda)
    idx = torch.zeros((B, n, 3), dtype=torch.int32, device=cuda)
    ops.three_nn_batch(B, n, m, u, k, d2, idx)
    for b in range(B):
        do, io = op.three_nn(unknown[b], np.array([n], np.int32), known[b], np.array([m], np.int32))
        assert np.array_equal(idx[b].cpu().numpy(), io)
        assert np.array_equal(d2[b].cpu().numpy(), do)
    feats = rng.normal(size=(B, C, m)).astype(np.float32)
    w = rng.uniform(0, 1, (B, n, 3)).astype(np.float32)
    w /= w.sum(-1, keepdims=True)
    f, wt = cu(feats, cuda), cu(w, cuda)
    out = torch.zeros((B, C, n), device=cuda)
    ops.three_interpolate_batch(B, C, m, n, f, idx, wt, out)
    ii = idx.cpu().numpy()
    expect = np.stack([sum(w[b, :, j][None, :].astype(np.float64) * feats[b][:, ii[b, :, j]] for j in range(3)) for b in range(B)])
    assert np.allclose(out.cpu().numpy(), expect, rtol=1e-5, atol=1e-6)
    g = rng.normal(size=(B, C, n)).astype(np.float32)
    gp = torch.zeros((B, C, m), device=cuda)
    ops.three_interpolate_grad_batch(B, C, n, m, cu(g, cuda), idx, wt, gp)
    eg = np.zeros((B, C, m), np.float64)
    for b in range(B):
        for j in range(3):
            for c in range(C):
                np.add.at(eg[b, c], ii[b, :, j], g[b, c].astype(np.float64) * w[b, :, j])
    assert np.allclose(gp.cpu().numpy(), eg, rtol=1e-4, atol=1e-5)
    ref = ref_batch_kernels()
    if ref is not None:
        rd, ri, ro = torch.zeros_like(d2), torch.zeros_like(idx), torch.zeros_like(out)
        torch.cuda.synchronize()
        ref.refb_three_nn(B, n, m, P(u), P(k), P(rd), P(ri))
        ref.refb_three_interpolate(B, C, m, n, P(f), P(ri), P(wt), P(ro))
        assert ref.refb_sync() == 0
        assert torch.equal(idx, ri) and torch.equal(d2, rd)
        assert torch.equal(out, ro)                                  # same contraction: bit-equal


def _voxel_table(rng, B, Z, Y, X, n_per_frame, voxel, origin):
    """Random occupied voxels with one representative point each -> xyz (N,3), table (B,Z,Y,X) of point rows (-1: empty)."""
    table = np.full((B, Z, Y, X), -1, np.int32)
    xyz = []
    for b in range(B):
        flat = rng.choice(Z * Y * X, n_per_frame, replace=False)
        z, y, x = np.unravel_index(flat, (Z, Y, X))
        centre = (np.stack([x, y, z], 1) + rng.uniform(0.05, 0.95, (n_per_frame, 3))) * voxel + origin
        table[b, z, y, x] = np.arange(n_per_frame) + b * n_per_frame
        xyz.append(centre.astype(np.float32))
    return np.concatenate(xyz), table


@pytest.mark.parametrize("rng_zyx,radius,nsample", [((1, 2, 2), 0.6, 16), ((2, 4, 4), 1.1, 16), ((0, 1, 1), 0.25, 4), ((3, 3, 3), 50.0, 32)])
def test_voxel_query(cuda, rng_zyx, radius, nsample):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(int(radius * 100) + nsample)
    B, Z, Y, X = 2, 10, 40, 44
    voxel, origin = np.array([0.4, 0.4, 0.5]), np.array([-8.0, -8.0, -3.0])
    xyz, table = _voxel_table(rng, B, Z, Y, X, 6000, voxel, origin)
    M = 1500
    new_xyz = np.concatenate([rng.uniform([-8.5, -8.5, -3.2], [9.6, 8.2, 2.2], (M - 100, 3)), xyz[:100].astype(np.float64) + 0.01]).astype(np.float32)
    zyx = np.floor((new_xyz[:, ::-1].astype(np.float64) - origin[::-1]) / voxel[::-1]).astype(np.int64)
    zyx = np.clip(zyx, [-1, -2, -2], [Z, Y + 1, X + 1])          # some centres just outside the grid: the bounds tests matter
    new_coords = np.concatenate([rng.integers(0, B, (M, 1)), zyx], 1).astype(np.int32)
    new_coords[-100:, 0] = 0
    idx = torch.zeros((M, nsample), dtype=torch.int32, device=cuda)
    a, c, d, e = cu(new_xyz, cuda), cu(xyz, cuda), cu(new_coords, cuda), cu(table, cuda)
    ops.voxel_query(M, Z, Y, X, nsample, radius, *rng_zyx, a, c, d, e, idx)
    o = op.voxel_query(rng_zyx, radius, nsample, xyz, new_xyz, new_coords, table)
    assert (o[:, 0] >= 0).any() and ((o[:, 0] == -1).any() or radius > 10)
    assert np.array_equal(idx.cpu().numpy(), o)
    ref = ref_batch_kernels()
    if ref is not None:
        r = torch.zeros_like(idx)
        torch.cuda.synchronize()
        ref.refb_voxel_query(M, Z, Y, X, nsample, ctypes.c_float(radius), *rng_zyx, P(a), P(c), P(d), P(e), P(r))
        assert ref.refb_sync() == 0
        assert torch.equal(idx, r)


@pytest.mark.parametrize("N,M,C,S", [(16384, 64, 16, 512), (3000, 100, 128, 64), (500, 7, 1, 512), (40, 3, 4, 16)])
def test_roipoint_pool3d(cuda, N, M, C, S):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(N + M + C + S)
    B = 2
    xyz = rng.uniform([-12, -12, -2.5], [12, 12, 1.5], (B, N, 3)).astype(np.float32)
    feat = rng.normal(size=(B, N, C)).astype(np.float32)
    boxes = np.stack([rand_boxes(rng, M, 10) for _ in range(B)])
    boxes[:, 0, :3] = 500.0                                        # a box without points
    boxes[0, 1, 3:6] = [30.0, 30.0, 8.0]                           # a box with more than S points
    x, bx, f = cu(xyz, cuda), cu(boxes, cuda), cu(feat, cuda)
    pooled = torch.zeros((B, M, S, 3 + C), device=cuda)
    empty = torch.zeros((B, M), dtype=torch.int32, device=cuda)
    ops.roipoint_pool3d_forward(x, bx, f, pooled, empty)
    e = empty.cpu().numpy()
    assert e[:, 0].all() and not e[:, 1].any()
    ref = ref_batch_kernels()
    if ref is not None:
        rp, re = torch.zeros_like(pooled), torch.zeros_like(empty)
        torch.cuda.synchronize()
        ref.refb_roipoint_pool3d(B, N, M, C, S, P(x), P(bx), P(f), P(rp), P(re))
        assert ref.refb_sync() == 0
        assert torch.equal(empty, re)
        assert torch.equal(pooled, rp)                             # bit-exact vs the reference kernels
    # C oracle: libm sinf/cosf may differ from CUDA's in the last ulp -> only boxes with a point on a face may differ
    bad = 0
    for b in range(B):
        po, eo = op.roipoint_pool3d(xyz[b], boxes[b], feat[b], S)
        assert np.array_equal(e[b], eo) or (e[b] != eo).sum() <= 1
        bad += int((np.abs(pooled[b].cpu().numpy() - po).reshape(M, -1).max(1) > 0).sum())
    assert bad <= max(1, (B * M) // 50)


def test_reference_batch_wrappers_over_dropin(cuda):
    """The reference's pointnet2_batch/pointnet2_utils.py, roipoint_pool3d_utils.py and voxel_query_utils.py imported as they
    are (autograd Functions, QueryAndGroup) with their compiled modules answered by this library."""
    if ref_env.install() is None:
        pytest.skip("no reference tree (neither /root/reference nor baseline/_ref)")
    pb = ref_env.ref("pcdet.ops.pointnet2.pointnet2_batch.pointnet2_utils")
    rp = ref_env.ref("pcdet.ops.roipoint_pool3d.roipoint_pool3d_utils")
    vq = ref_env.ref("pcdet.ops.pointnet2.pointnet2_stack.voxel_query_utils")
    assert pb.pointnet2.__name__ == "pcdet_ops.pointnet2_batch_cuda"
    assert rp.roipoint_pool3d_cuda.__name__ == "pcdet_ops.roipoint_pool3d_cuda"
    from oracle import pointnet2 as op
    rng = np.random.default_rng(5)
    B, n, m, C = 2, 2048, 256, 8
    xyz = np.stack([_cloud(rng, n) for _ in range(B)])
    x = cu(xyz, cuda)
    fps = pb.furthest_point_sample(x, m)
    for b in range(B):
        assert np.array_equal(fps[b].cpu().numpy(), op.farthest_point_sampling(xyz[b], m)[0])
    feats = torch.randn(B, C, n, device=cuda, requires_grad=True)
    new_xyz = pb.gather_operation(x.transpose(1, 2).contiguous(), fps).transpose(1, 2).contiguous()
    assert torch.equal(new_xyz, torch.gather(x, 1, fps.long()[..., None].expand(-1, -1, 3)))
    grouper = pb.QueryAndGroup(0.9, 16, use_xyz=True)
    g = grouper(x, new_xyz, feats)                                 # (B, 3 + C, m, 16)
    assert g.shape == (B, 3 + C, m, 16)
    idx = pb.ball_query(0.9, 16, x, new_xyz).long()
    expect_f = torch.gather(feats.unsqueeze(2).expand(-1, -1, m, -1), 3, idx.unsqueeze(1).expand(-1, C, -1, -1))
    assert torch.equal(g[:, 3:], expect_f)
    # backward through grouping_operation == backward through torch.gather
    gr = torch.randn_like(g)
    (ga,) = torch.autograd.grad(g, feats, gr, retain_graph=True)
    (gb,) = torch.autograd.grad(expect_f, feats, gr[:, 3:])
    assert torch.allclose(ga, gb, rtol=1e-4, atol=1e-5)
    # feature propagation: three_nn + three_interpolate (+ grad)
    dist, i3 = pb.three_nn(x, new_xyz)
    w = 1.0 / (dist + 1e-8)
    w = w / w.sum(2, keepdim=True)
    kf = torch.randn(B, C, m, device=cuda, requires_grad=True)
    interp = pb.three_interpolate(kf, i3, w)
    expect_i = sum(w[:, None, :, j] * torch.gather(kf, 2, i3[:, None, :, j].long().expand(-1, C, -1)) for j in range(3))
    assert torch.allclose(interp, expect_i, rtol=1e-5, atol=1e-6)
    go = torch.randn_like(interp)
    (g1,) = torch.autograd.grad(interp, kf, go, retain_graph=True)
    (g2,) = torch.autograd.grad(expect_i, kf, go)
    assert torch.allclose(g1, g2, rtol=1e-4, atol=1e-5)
    # RoIPointPool3d module (enlarges the boxes itself, box_utils.enlarge_box3d)
    boxes = cu(np.stack([rand_boxes(rng, 20, 7) for _ in range(B)]), cuda)
    pf = torch.randn(B, n, C, device=cuda)
    pooled, empty = rp.RoIPointPool3d(num_sampled_points=128, pool_extra_width=(1.0, 1.0, 1.0))(x, pf, boxes)
    assert pooled.shape == (B, 20, 128, 3 + C) and empty.shape == (B, 20)
    big = boxes.clone()
    big[..., 3:6] += 1.0
    for b in range(B):
        po, eo = op.roipoint_pool3d(xyz[b], big[b].cpu().numpy(), pf[b].cpu().numpy(), 128)
        assert (empty[b].cpu().numpy() != eo).sum() <= 1
        assert (np.abs(pooled[b].cpu().numpy() - po).reshape(20, -1).max(1) > 0).sum() <= 1
    # VoxelQuery function: empty mask + idx with the -1 rows zeroed (voxel_query_utils.py:33-42)
    Z, Y, X = 8, 30, 30
    vx, table = _voxel_table(rng, 1, Z, Y, X, 3000, np.array([0.5, 0.5, 0.5]), np.array([-7.5, -7.5, -2.0]))
    q = vx[:400] + np.float32(0.02)
    qc = np.concatenate([np.zeros((400, 1)), np.floor((q[:, ::-1].astype(np.float64) - np.array([-2.0, -7.5, -7.5])) / 0.5)], 1).astype(np.int32)
    qc[:, 1:] = np.clip(qc[:, 1:], 0, [Z - 1, Y - 1, X - 1])
    vi, vempty = vq.voxel_query((1, 2, 2), 0.7, 8, cu(vx, cuda), cu(q, cuda), cu(qc, cuda), cu(table, cuda))
    o = op.voxel_query((1, 2, 2), 0.7, 8, vx, q, qc, table)
    assert np.array_equal(vempty.cpu().numpy(), o[:, 0] == -1)
    o[o[:, 0] == -1] = 0
    assert np.array_equal(vi.cpu().numpy(), o)


# ---------------------------------------------------------------------------------------------- vector pool (PV-RCNN++)
def _vp_inputs(rng, cuda, radius):
    xyz_cnt = np.array([4000, 2500], np.int32)
    new_cnt = np.array([300, 211], np.int32)
    xyz = np.concatenate([_cloud(rng, 4000, 6.0), _cloud(rng, 2500, 6.0)])
    new_xyz = np.concatenate([xyz[:300] + np.float32(0.05), xyz[4000:4211] + np.float32(0.02)])
    new_xyz[5] = [50, 50, 50]                                        # a query without neighbours
    ang = rng.uniform(0, 2 * np.pi, 32)                               # support points at distance == radius of query 0
    xyz[100:132] = new_xyz[0] + np.float32(radius) * np.stack([np.cos(ang), np.sin(ang), np.zeros(32)], 1).astype(np.float32)
    return xyz, xyz_cnt, new_xyz, new_cnt


def _lists(stack, start_len):
    s, sl = stack.cpu().numpy(), start_len.cpu().numpy()
    return [s[o:o + n].tolist() for o, n in sl]


@pytest.mark.parametrize("neighbor_type,nsample,radius", [(1, -1, 0.8), (0, -1, 0.6), (1, 16, 1.5), (0, 7, 3.0)])
def test_stacked_local_neighbors_and_three_nn(cuda, neighbor_type, nsample, radius):
    """Per-query neighbour lists (content and order) and the 3-NN of every grid centre == the reference kernels; the reference
    places the lists with a global atomicAdd (layout depends on scheduling), this library in query order."""
    from crb3d import ops
    rng = np.random.default_rng(int(radius * 10) + neighbor_type)
    xyz, xyz_cnt, new_xyz, new_cnt = _vp_inputs(rng, cuda, radius)
    M = len(new_xyz)
    x, xc, nx, nc = cu(xyz, cuda), cu(xyz_cnt, cuda), cu(new_xyz, cuda), cu(new_cnt, cuda)
    avg = 1000
    stack = torch.zeros(avg * M, dtype=torch.int32, device=cuda)
    start_len = torch.zeros((M, 2), dtype=torch.int32, device=cuda)
    cumsum = torch.zeros(1, dtype=torch.int32, device=cuda)
    ops.query_stacked_local_neighbor_idxs(x, xc, nx, nc, stack, start_len, cumsum, avg, radius, nsample, neighbor_type)
    mine = _lists(stack, start_len)
    sl = start_len.cpu().numpy()
    assert int(cumsum.item()) == int(sl[:, 1].sum()) and np.array_equal(sl[:, 0], np.concatenate([[0], np.cumsum(sl[:, 1])[:-1]]))
    assert len(mine[5]) == 0 and max(len(l) for l in mine) > 3 and (nsample < 0 or max(len(l) for l in mine) == nsample)
    # brute force in float64 away from the boundary: every list holds frame-local neighbours in ascending order
    for q in (0, 17, 299, 300, 510):
        f0 = 0 if q < 300 else 4000
        assert mine[q] == sorted(mine[q]) and all(f0 <= i < f0 + (4000 if q < 300 else 2500) for i in mine[q])
    G = 8
    centers = (new_xyz[:, None, :] + rng.uniform(-radius, radius, (M, G, 3))).astype(np.float32)
    ct = cu(centers, cuda)
    idxs = torch.full((M, G, 3), -1, dtype=torch.int32, device=cuda)
    d2 = torch.zeros((M, G, 3), device=cuda)
    ops.query_three_nn_by_stacked_local_idxs(x, nx, ct, idxs, d2, stack[: int(cumsum.item())], start_len, M, G)
    assert int((idxs[5] == -1).all()) == 1
    ref = ref_batch_kernels()
    if ref is not None:
        rstack, rsl, rcs = torch.zeros_like(stack), torch.zeros_like(start_len), torch.zeros_like(cumsum)
        torch.cuda.synchronize()
        ref.refb_local_neighbors(P(x), P(xc), P(nx), P(nc), P(rstack), P(rsl), P(rcs), avg, ctypes.c_float(radius), 2, M, nsample, neighbor_type)
        assert ref.refb_sync() == 0
        assert int(rcs.item()) == int(cumsum.item())
        assert _lists(rstack, rsl) == mine                            # same neighbours in the same order, query by query
        ridx, rd2 = torch.full_like(idxs, -1), torch.zeros_like(d2)
        ref.refb_three_nn_local(P(x), P(nx), P(ct), P(ridx), P(rd2), P(rstack), P(rsl), M, G)
        assert ref.refb_sync() == 0
        assert torch.equal(idxs, ridx) and torch.equal(d2, rd2)


@pytest.mark.parametrize("pooling_type,neighbor_type,nsample,c_in,ce", [(0, 0, -1, 32, 16), (0, 1, 24, 16, 16), (1, 0, -1, 32, 8), (0, 0, -1, 5, 5)])
def test_vector_pool_forward_and_grad(cuda, pooling_type, neighbor_type, nsample, c_in, ce):
    from crb3d import ops
    rng = np.random.default_rng(c_in * 3 + pooling_type + nsample % 7)
    radius = 1.2
    xyz, xyz_cnt, new_xyz, new_cnt = _vp_inputs(rng, cuda, radius)
    N, M, (gx, gy, gz) = len(xyz), len(new_xyz), (3, 3, 2)
    G, c_out = gx * gy * gz, ce * gx * gy * gz
    feat = rng.normal(size=(N, c_in)).astype(np.float32)
    x, xc, ft, nx, nc = cu(xyz, cuda), cu(xyz_cnt, cuda), cu(feat, cuda), cu(new_xyz, cuda), cu(new_cnt, cuda)
    max_sum = 400 * M

    def run(fn_is_ref):
        nf = torch.zeros((M, c_out), device=cuda)
        nl = torch.zeros((M, 3 * G), device=cuda)
        pc = torch.zeros((M, G), dtype=torch.int32, device=cuda)
        gi = torch.zeros((max_sum, 3), dtype=torch.int32, device=cuda)
        if fn_is_ref:
            torch.cuda.synchronize()
            n = ref_batch_kernels().refb_vector_pool(P(x), P(ft), P(xc), P(nx), P(nf), P(nl), P(nc), P(pc), P(gi), gx, gy, gz, ctypes.c_float(radius),
                                                     2, N, M, c_in, c_out, G, 1, max_sum, nsample, neighbor_type, pooling_type)
        else:
            n = ops.vector_pool(x, xc, ft, nx, nc, nf, nl, pc, gi, gx, gy, gz, radius, 1, max_sum, nsample, neighbor_type, pooling_type)
        return nf, nl, pc, gi[:n], n

    nf, nl, pc, gi, n = run(False)
    assert 0 < n <= max_sum and int(pc.sum()) >= n and float(nf.abs().max()) > 0 and int(pc[5].sum()) == 0
    # grouped rows: (support point, query, sub-voxel); avg pooling: the sub-voxel sums are the sums of the grouped points' features
    g = gi.cpu().numpy()
    assert (g[:, 1] >= 0).all() and (g[:, 1] < M).all() and (g[:, 2] < G).all()
    if pooling_type == 0 and nsample < 0:
        expect = np.zeros((M, G, ce), np.float64)
        np.add.at(expect, (g[:, 1], g[:, 2]), feat[g[:, 0]].reshape(len(g), c_in // ce, ce).sum(1).astype(np.float64))
        assert np.allclose(nf.cpu().numpy().reshape(M, G, ce), expect, rtol=1e-4, atol=1e-4)
        assert np.array_equal(np.bincount(g[:, 1] * G + g[:, 2], minlength=M * G).reshape(M, G), pc.cpu().numpy())
    go = torch.randn(M, c_out, device=cuda)
    gs = torch.zeros((N, c_in), device=cuda)
    ops.vector_pool_grad(go, pc, gi.contiguous(), gs)
    if ref_batch_kernels() is not None:
        rf, rl, rp, rg, rn = run(True)
        assert rn == n
        assert torch.equal(pc, rp) and torch.equal(nf, rf) and torch.equal(nl, rl)       # sums accumulated in the same order: bit-equal
        key = lambda t: np.array(sorted(map(tuple, t.cpu().numpy().tolist())))
        assert np.array_equal(key(gi), key(rg))                                           # same grouped set (the order is scheduling)
        rgs = torch.zeros_like(gs)
        torch.cuda.synchronize()
        ref_batch_kernels().refb_vector_pool_grad(P(go), P(rp), P(rg.contiguous()), P(rgs), N, M, c_out, c_in, G, rn)
        assert ref_batch_kernels().refb_sync() == 0
        assert torch.allclose(gs, rgs, rtol=1e-5, atol=1e-6)


def test_reference_vector_pool_functions_over_dropin(cuda):
    """The reference's VectorPoolWithVoxelQuery / ThreeNNForVectorPoolByTwoStep autograd functions (pointnet2_stack/pointnet2_utils.py:
    302-448), unmodified, with every pointnet2_stack_cuda entry point answered by this library."""
    if ref_env.install() is None:
        pytest.skip("no reference tree (neither /root/reference nor baseline/_ref)")
    pu = ref_env.ref("pcdet.ops.pointnet2.pointnet2_stack.pointnet2_utils")
    rng = np.random.default_rng(9)
    xyz, xyz_cnt, new_xyz, new_cnt = _vp_inputs(rng, cuda, 1.0)
    x, xc, nx, nc = cu(xyz, cuda), cu(xyz_cnt, cuda), cu(new_xyz, cuda), cu(new_cnt, cuda)
    feat = torch.randn(len(xyz), 32, device=cuda, requires_grad=True)
    new_feat, new_local, n_mean, pcnt = pu.vector_pool_with_voxel_query_op(x, xc, feat, nx, nc, 3, 3, 3, 1.0, 16, True, 2, -1, 0, 0)
    assert new_feat.shape == (len(new_xyz), 16 * 27) and pcnt.shape == (len(new_xyz), 27) and int(pcnt.sum()) > 0
    (g,) = torch.autograd.grad(new_feat.sum(), feat)
    # every grouped support point receives 1 / (points in its sub-voxel) per output channel it feeds
    assert float(g.abs().sum()) > 0 and torch.isfinite(g).all()
    centers = (new_xyz[:, None, :] + rng.uniform(-1, 1, (len(new_xyz), 27, 3))).astype(np.float32)
    dist, idx, avg = pu.three_nn_for_vector_pool_by_two_step(x, xc, nx, cu(centers, cuda), nc, 1.0, -1, 1, 5, 27, 2.0)
    assert dist.shape == idx.shape == (len(new_xyz), 27, 3) and int((idx[5] == -1).all()) == 1 and int((idx[0] >= 0).all()) == 1
    # nearest of the local list == brute-force nearest within the query radius (2.0) in the frame
    q = 0
    d = ((xyz[:4000][None] - centers[q][:, None]) ** 2).sum(-1)
    inside = ((xyz[:4000] - new_xyz[q]) ** 2).sum(-1) <= 4.0 - 1e-4
    d[:, ~inside] = np.inf
    assert np.array_equal(idx[q, :, 0].cpu().numpy(), d.argmin(1))


def test_extra_ops_empty_inputs(cuda):
    """Zero-sized problems return without a launch and leave the caller's (zero-filled) outputs untouched."""
    from crb3d import ops
    z3 = torch.zeros((0, 3), device=cuda)
    idx = torch.zeros((0, 8), dtype=torch.int32, device=cuda)
    ops.voxel_query(0, 4, 4, 4, 8, 1.0, 1, 1, 1, z3, torch.zeros((5, 3), device=cuda), torch.zeros((0, 4), dtype=torch.int32, device=cuda),
                    torch.full((1, 4, 4, 4), -1, dtype=torch.int32, device=cuda), idx)
    bidx = torch.zeros((2, 0, 4), dtype=torch.int32, device=cuda)
    ops.ball_query_batch(2, 10, 0, 1.0, 4, torch.zeros((2, 0, 3), device=cuda), torch.zeros((2, 10, 3), device=cuda), bidx)
    pooled = torch.zeros((1, 0, 16, 5), device=cuda)
    ops.roipoint_pool3d_forward(torch.zeros((1, 7, 3), device=cuda), torch.zeros((1, 0, 7), device=cuda), torch.zeros((1, 7, 2), device=cuda),
                                pooled, torch.zeros((1, 0), dtype=torch.int32, device=cuda))
    # a frame without points: every box is flagged empty
    ef = torch.zeros((1, 3), dtype=torch.int32, device=cuda)
    ops.roipoint_pool3d_forward(torch.zeros((1, 0, 3), device=cuda), torch.ones((1, 3, 7), device=cuda), torch.zeros((1, 0, 2), device=cuda),
                                torch.zeros((1, 3, 4, 5), device=cuda), ef)
    assert ef.tolist() == [[1, 1, 1]]
    torch.cuda.synchronize()
