"""-m gpu parity: ball query / grouping / FPS / 3-NN / interpolation - ours vs the C oracle (bit-exact: no libm
transcendental is involved) and vs the reference's own CUDA kernels."""
import ctypes

import numpy as np
import pytest
import torch

from util import P, cu, ref_kernels

pytestmark = pytest.mark.gpu


def _cloud(rng, n, spread=20.0):
    p = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    p[:, 2] = rng.uniform(-2, 1, n)
    return p


@pytest.mark.parametrize("radius,nsample", [(0.4, 16), (0.8, 16), (2.4, 32), (4.8, 32)])
def test_ball_query(cuda, radius, nsample):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(int(radius * 10) + nsample)
    xyz_cnt = np.array([9000, 1, 7000], np.int32)
    new_cnt = np.array([700, 300, 513], np.int32)
    xyz = _cloud(rng, int(xyz_cnt.sum()), 8.0)
    new_xyz = np.concatenate([xyz[:700] + np.float32(0.05), _cloud(rng, 300, 8.0), xyz[9001:9514] + np.float32(0.01)])
    # adversarial: sources at distance == radius (within float rounding) from query 0
    q0 = new_xyz[0]
    ang = rng.uniform(0, 2 * np.pi, 64)
    xyz[100:164] = q0 + np.float32(radius) * np.stack([np.cos(ang), np.sin(ang), np.zeros(64)], 1).astype(np.float32)
    idx = torch.zeros((len(new_xyz), nsample), dtype=torch.int32, device=cuda)
    ops.ball_query(3, len(new_xyz), radius, nsample, cu(new_xyz, cuda), cu(new_cnt, cuda), cu(xyz, cuda), cu(xyz_cnt, cuda), idx)
    ref_o = op.ball_query(radius, nsample, xyz, xyz_cnt, new_xyz, new_cnt)
    assert np.array_equal(idx.cpu().numpy(), ref_o)
    assert (ref_o[:, 0] == -1).any() and (ref_o[:, 0] >= 0).any()
    ref = ref_kernels()
    if ref is not None:
        r = torch.zeros_like(idx)
        a, b, c, d = cu(new_xyz, cuda), cu(new_cnt, cuda), cu(xyz, cuda), cu(xyz_cnt, cuda)
        torch.cuda.synchronize()
        ref.ref_ball_query(3, len(new_xyz), ctypes.c_float(radius), nsample, P(a), P(b), P(c), P(d), P(r))
        assert ref.ref_sync() == 0
        assert torch.equal(idx, r)


@pytest.mark.parametrize("C,ns", [(1, 16), (16, 16), (64, 32), (128, 16)])
def test_group_points_and_grad(cuda, C, ns):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(C + ns)
    feat_cnt = np.array([3000, 2000], np.int32)
    idx_cnt = np.array([400, 250], np.int32)
    feat = rng.normal(size=(5000, C)).astype(np.float32)
    idx = np.concatenate([rng.integers(0, 3000, (400, ns)), rng.integers(0, 2000, (250, ns))]).astype(np.int32)
    out = torch.zeros((650, C, ns), device=cuda)
    ops.group_points(2, 650, C, ns, cu(feat, cuda), cu(feat_cnt, cuda), cu(idx, cuda), cu(idx_cnt, cuda), out)
    assert np.array_equal(out.cpu().numpy(), op.group_points(feat, feat_cnt, idx, idx_cnt))
    g = rng.normal(size=(650, C, ns)).astype(np.float32)
    gf = torch.zeros((5000, C), device=cuda)
    ops.group_points_grad(2, 650, C, 5000, ns, cu(g, cuda), cu(idx, cuda), cu(idx_cnt, cuda), cu(feat_cnt, cuda), gf)
    assert np.allclose(gf.cpu().numpy(), op.group_points_grad(g, idx, idx_cnt, feat_cnt, 5000), rtol=1e-4, atol=1e-5)


@pytest.mark.parametrize("n,m", [(20000, 2048), (40000, 2048), (65000, 300), (70000, 64), (16384, 1024), (5000, 512), (1024, 256), (1000, 128), (700, 700), (33, 8), (1, 1)])
def test_farthest_point_sampling(cuda, n, m):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(n + m)
    pts = np.stack([_cloud(rng, n), _cloud(rng, n)])
    if n >= 700:
        pts[0, 50:60] = pts[0, 40:50]     # exact duplicates -> exact distance ties
    temp = torch.full((2, n), 1e10, device=cuda)
    idx = torch.zeros((2, m), dtype=torch.int32, device=cuda)
    ops.farthest_point_sampling(2, n, m, cu(pts, cuda), temp, idx)
    for b in range(2):
        io, to = op.farthest_point_sampling(pts[b], m)
        assert np.array_equal(idx[b].cpu().numpy(), io)
        if m > 1:
            assert np.array_equal(temp[b].cpu().numpy(), to)
    ref = ref_kernels()
    if ref is not None:
        rt = torch.full((2, n), 1e10, device=cuda)
        ri = torch.zeros((2, m), dtype=torch.int32, device=cuda)
        pc = cu(pts, cuda)
        torch.cuda.synchronize()
        ref.ref_fps(2, n, m, P(pc), P(rt), P(ri))
        assert ref.ref_sync() == 0
        assert torch.equal(idx, ri)


def test_stack_farthest_point_sampling(cuda):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(77)
    cnt = np.array([6000, 1500, 9000], np.int32)
    ms = np.array([512, 100, 700], np.int32)
    pts = _cloud(rng, int(cnt.sum()))
    temp = torch.full((int(cnt.sum()),), 1e10, device=cuda)
    idx = torch.zeros((int(ms.sum()),), dtype=torch.int32, device=cuda)
    ops.stack_farthest_point_sampling(cu(pts, cuda), temp, cu(cnt, cuda), idx, cu(ms, cuda))
    s = o = 0
    for b in range(3):
        io, _ = op.farthest_point_sampling(pts[s:s + cnt[b]], int(ms[b]), block=1024)
        assert np.array_equal(idx[o:o + ms[b]].cpu().numpy(), io + s)
        s += cnt[b]; o += ms[b]


def test_three_nn_interpolate(cuda):
    from crb3d import ops
    from oracle import pointnet2 as op
    rng = np.random.default_rng(55)
    uc, kc = np.array([4000, 2500], np.int32), np.array([600, 2], np.int32)
    unknown, known = _cloud(rng, 6500), _cloud(rng, 602)
    d2 = torch.zeros((6500, 3), device=cuda)
    idx = torch.zeros((6500, 3), dtype=torch.int32, device=cuda)
    ops.three_nn(2, 6500, 602, cu(unknown, cuda), cu(uc, cuda), cu(known, cuda), cu(kc, cuda), d2, idx)
    d_o, i_o = op.three_nn(unknown, uc, known, kc)
    assert np.array_equal(idx.cpu().numpy(), i_o) and np.array_equal(d2.cpu().numpy(), d_o)
    feat = rng.normal(size=(602, 32)).astype(np.float32)
    w = rng.uniform(0, 1, (6500, 3)).astype(np.float32)
    out = torch.zeros((6500, 32), device=cuda)
    ops.three_interpolate(6500, 32, cu(feat, cuda), idx, cu(w, cuda), out)
    assert np.allclose(out.cpu().numpy(), op.three_interpolate(feat, i_o, w), rtol=1e-5, atol=1e-6)
    g = rng.normal(size=(6500, 32)).astype(np.float32)
    gf = torch.zeros((602, 32), device=cuda)
    ops.three_interpolate_grad(6500, 32, cu(g, cuda), idx, cu(w, cuda), gf)
    ref = np.zeros((602, 32), np.float64)
    for j in range(3):
        np.add.at(ref, i_o[:, j], g.astype(np.float64) * w[:, j:j + 1])
    assert np.allclose(gf.cpu().numpy(), ref, rtol=1e-3, atol=1e-3)
