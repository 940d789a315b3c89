"""-m gpu: size-independent properties of the hot-path kernels at BASELINE.json's FULL sizes (configs[1]: 16 KITTI-shaped frames,
~20 k points each, 1408 x 1600 x 40 voxel grid), where the CPU oracle would take minutes. The small-size parity tests pin the
values; these pin that nothing changes with size: set equality under permutation, rulebook soundness / completeness / mirror
symmetry checked against the coordinates themselves, conv linearity and exact gather through a one-hot kernel, NMS invariants,
FPS monotonicity, dense round trips. Every check runs on the device with plain torch ops on the kernels' outputs."""
import numpy as np
import pytest
import torch

from util import cu, rand_boxes

pytestmark = pytest.mark.gpu

K_RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
K_VOX = [0.05, 0.05, 0.1]
B = 16


def _pool(dev, shuffle_seed=None):
    from crb3d import synth
    frames = [synth.make_frame(i) for i in range(B)]
    if shuffle_seed is not None:
        rng = np.random.default_rng(shuffle_seed)
        frames = [f[rng.permutation(len(f))] for f in frames]
    offs = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    return cu(np.concatenate(frames), dev), cu(offs, dev)


def _keys(c, shape):
    c = c.long()
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


@pytest.fixture(scope="module")
def geom(cuda):
    """Voxels and the 8 rulebooks of the full batch, built once."""
    from crb3d import second
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(cuda)
    pts, offs = _pool(cuda)
    g = model.geometry(pts, offs, B)
    return model, pts, offs, g


def test_voxelize_full_batch_invariants_and_permutation(cuda, geom):
    from crb3d import ops
    model, pts, offs, g = geom
    res = ops.voxelize(pts, offs, B, K_RANGE, K_VOX, 5, 40000)
    c, num, mean, voff = res["coords"], res["num_points"], res["mean"], res["frame_voxel_offsets"]
    grid = (40, 1600, 1408)
    m = c.shape[0]
    assert 200000 < m <= B * 40000 and int(voff[-1]) == m and int(voff[0]) == 0
    assert bool((voff[1:] >= voff[:-1]).all())
    # rows are grouped by frame in frame order, inside the grid, and no cell occurs twice
    assert bool((c[1:, 0] >= c[:-1, 0]).all())
    assert torch.equal(torch.bincount(c[:, 0].long(), minlength=B).int(), voff[1:] - voff[:-1])
    for j, g_j in enumerate(grid):
        assert int(c[:, 1 + j].min()) >= 0 and int(c[:, 1 + j].max()) < g_j
    keys = _keys(c, grid)
    assert torch.unique(keys).numel() == m
    assert int(num.min()) >= 1 and int(num.max()) <= 5
    # the mean of a voxel's points lies inside the voxel (cell bounds computed in float64; 1e-4 m of slack for the float32 floor)
    lo = torch.tensor(K_RANGE[:3], dtype=torch.float64, device=cuda)
    vs = torch.tensor(K_VOX, dtype=torch.float64, device=cuda)
    cell_lo = lo + c[:, [3, 2, 1]].double() * vs
    xyz = mean[:, :3].double()
    assert bool(((xyz >= cell_lo - 1e-4) & (xyz <= cell_lo + vs + 1e-4)).all())
    # points shuffled inside their frames: the same SET of voxels with the same counts; voxels below the 5-point cap (where
    # every point is kept whatever the order) have the same mean up to the summation order
    pts2, offs2 = _pool(cuda, shuffle_seed=11)
    res2 = ops.voxelize(pts2, offs2, B, K_RANGE, K_VOX, 5, 40000)
    k2 = _keys(res2["coords"], grid)
    o1, o2 = torch.argsort(keys), torch.argsort(k2)
    assert torch.equal(keys[o1], k2[o2])
    assert torch.equal(num[o1], res2["num_points"][o2])
    below = num[o1] < 5
    d = (mean[o1][below] - res2["mean"][o2][below]).abs().max()
    assert float(d) < 1e-4, float(d)


def _offsets(ksize, dev):
    kz, ky, kx = torch.meshgrid(torch.arange(ksize[0]), torch.arange(ksize[1]), torch.arange(ksize[2]), indexing="ij")
    return torch.stack([kz.reshape(-1), ky.reshape(-1), kx.reshape(-1)], 1).to(dev)       # k = (kz * Ky + ky) * Kx + kx


def _check_book(d, batch):
    """Soundness: every table entry joins an (input, output) pair whose coordinates satisfy in = out * stride - pad + k * dil.
    Completeness: the number of entries equals the number of such pairs found by an independent sorted-key search."""
    dev = d.nbr.device
    oc, ic = d.out_indices.long(), d.indices.long()
    K, n_out = d.nbr.shape
    ks, st, pd, dl = list(d.ksize), list(d.stride), list(d.padding), list(d.dilation)
    offs = _offsets(ks, dev)
    assert offs.shape[0] == K
    in_shape = [int(s) for s in d.spatial_shape]
    in_keys = _keys(ic, in_shape)
    sk, order = torch.sort(in_keys)
    st_t = torch.tensor(st, device=dev)
    pd_t = torch.tensor(pd, device=dev)
    dl_t = torch.tensor(dl, device=dev)
    shape_t = torch.tensor(in_shape, device=dev)
    n_valid = 0
    for k in range(K):
        want = oc[:, 1:] * st_t - pd_t + offs[k] * dl_t                        # input cell each output row reads through offset k
        inside = ((want >= 0) & (want < shape_t)).all(1)
        wkey = _keys(torch.cat([oc[:, :1], want.clamp(min=0)], 1), in_shape)
        pos = torch.searchsorted(sk, wkey).clamp(max=sk.numel() - 1)
        found = inside & (sk[pos] == wkey)
        expect = torch.where(found, order[pos], torch.full_like(pos, -1))
        assert torch.equal(d.nbr[k].long(), expect), "offset %d" % k
        n_valid += int(found.sum())
    assert n_valid > 0
    if d.is_subm:
        # mirror symmetry: i reads j through offset k  <=>  j reads i through offset K-1-k; the centre reads itself
        rows = torch.arange(n_out, device=dev)
        assert torch.equal(d.nbr[K // 2].long(), rows)
        for k in range(K // 2):
            j = d.nbr[k].long()
            v = j >= 0
            assert torch.equal(d.nbr[K - 1 - k].long()[j[v]], rows[v])
    else:
        # output rows: ascending cell order, unique, each reached by at least one input; the transposed table is the inverse map
        ok = _keys(oc, [int(s) for s in d.out_spatial_shape])
        assert bool((ok[1:] > ok[:-1]).all())
        assert bool((d.nbr >= 0).any(0).all())
        assert int((d.nbr_t >= 0).sum()) == n_valid and bool((d.nbr_t >= 0).any(0).all())
        for k in range(K):
            i = d.nbr[k].long()
            v = i >= 0
            assert torch.equal(d.nbr_t[k].long()[i[v]], torch.arange(n_out, device=dev)[v])
    return n_valid


def test_rulebooks_full_batch_sound_complete_symmetric(cuda, geom):
    model, pts, offs, g = geom
    books = g["rulebooks"]
    seen = 0
    for key, d in books.items():
        if key.startswith("_"):
            continue
        n = _check_book(d, B)
        seen += 1
        assert n >= d.nbr.shape[1]
    assert seen == 8


def test_sparse_conv_full_batch_one_hot_gather_and_linearity(cuda, geom):
    """A kernel that is the identity at ONE offset turns the conv into a gather through that offset's table column: bit-exact
    at full size on the tensor-core path (TF32-representable inputs times 1.0, summed with zeros). Linearity within the TF32
    input-rounding bound."""
    from crb3d import ops
    model, pts, offs, g = geom
    books = g["rulebooks"]
    gen = torch.Generator(device="cpu").manual_seed(5)
    for key, c in (("subm3", 64), ("spconv4", 64), ("subm2", 32)):
        d = books[key]
        K, n_out = d.nbr.shape
        n_in = d.indices.shape[0]
        x = ops.round_tf32(torch.randn((n_in, c), generator=gen).to(cuda))
        for k in (0, K // 2, K - 1, 7):
            w = torch.zeros((c, K, c), device=cuda)
            w[:, k, :] = torch.eye(c, device=cuda)
            y = ops.spconv_forward(x, d.nbr, w, tf32=True)
            idx = d.nbr[k].long()
            ref = torch.where((idx >= 0).unsqueeze(1), x[idx.clamp(min=0)], torch.zeros((), device=cuda))
            assert torch.equal(y, ref), (key, k)
        w = ops.round_tf32(torch.randn((c, K, c), generator=gen).to(cuda) * 0.05)
        x2 = ops.round_tf32(torch.randn((n_in, c), generator=gen).to(cuda))
        y1, y2 = ops.spconv_forward(x, d.nbr, w, tf32=True), ops.spconv_forward(x2, d.nbr, w, tf32=True)
        y12 = ops.spconv_forward(ops.round_tf32(x + x2), d.nbr, w, tf32=True)      # the sum is rounded to TF32: 2^-11 relative per input
        rms = float(y12.pow(2).mean().sqrt())
        assert float((y12 - (y1 + y2)).abs().max()) < 4e-3 * rms
        # the exact-fp32 kernel is linear to fp32 rounding
        e1, e2 = ops.spconv_forward(x, d.nbr, w, tf32=False), ops.spconv_forward(x2, d.nbr, w, tf32=False)
        e12 = ops.spconv_forward(x + x2, d.nbr, w, tf32=False)
        assert float((e12 - (e1 + e2)).abs().max()) < 1e-4 * max(rms, 1.0)
        # and the two kernels agree on TF32-representable operands up to the summation order
        assert float((y1 - e1).abs().max()) < 2e-4 * max(rms, 1.0)


def test_dense_round_trip_full_batch(cuda, geom):
    from crb3d import ops
    model, pts, offs, g = geom
    d = g["rulebooks"]["spconv_down2"]
    oc, shape = d.out_indices, [int(s) for s in d.out_spatial_shape]
    n = oc.shape[0]
    feat = torch.randn((n, 128), device=cuda)
    for cl in (False, True):
        dense = ops.sparse_to_dense(feat, oc, B, shape, channels_last_bev=cl)
        assert int((dense != 0).sum()) == int((feat != 0).sum())
        assert torch.equal(ops.dense_to_sparse(dense, oc, 128, shape, channels_last_bev=cl), feat)
    a = ops.sparse_to_dense(feat, oc, B, shape, channels_last_bev=True).permute(0, 3, 1, 2)
    b = ops.sparse_to_dense(feat, oc, B, shape).view(B, 128 * shape[0], shape[1], shape[2])
    assert torch.equal(a, b)


def test_nms_invariants_at_pre_maxsize(cuda):
    """15 000 piled-up boxes (beyond NMS_PRE_MAXSIZE = 4096 of the KITTI configs): kept boxes do not overlap above the threshold,
    every dropped box is covered by an earlier kept one, and NMS of the kept list keeps everything."""
    from crb3d import ops
    rng = np.random.default_rng(2)
    n, thr = 15000, 0.1
    boxes = cu(rand_boxes(rng, n, 40, cluster=True), cuda)
    keep, num = ops.nms_sorted(boxes, thr)
    nk = int(num.item())
    keep = keep[:nk]
    assert 100 < nk < n and bool((keep[1:] > keep[:-1]).all())
    kept = boxes[keep]
    iou_kk = ops.boxes_iou_bev(kept, kept)
    assert float(torch.triu(iou_kk, 1).max()) <= thr
    iou_ak = ops.boxes_iou_bev(kept, boxes).t()                                 # (n, nk), argument order as the NMS kernel has it
    earlier = keep.unsqueeze(0) < torch.arange(n, device=cuda).unsqueeze(1)     # kept box ranks above the row's box
    covered = ((iou_ak > thr) & earlier).any(1)
    is_kept = torch.zeros(n, dtype=torch.bool, device=cuda)
    is_kept[keep] = True
    assert bool((covered | is_kept).all()) and not bool((covered & is_kept).any())
    keep2, num2 = ops.nms_sorted(kept, thr)
    assert int(num2.item()) == nk and torch.equal(keep2[:nk], torch.arange(nk, device=cuda))


def test_fps_full_size_monotone(cuda):
    """2048 keypoints out of 16 384 points per frame (PV-RCNN's VSA): distinct indices starting at 0, and the distance of each
    new sample to the samples before it never grows."""
    from crb3d import ops
    rng = np.random.default_rng(4)
    b, n, m = 4, 16384, 2048
    pts = cu(rng.uniform([0, -40, -3], [70, 40, 1], (b, n, 3)).astype(np.float32), cuda)
    temp = torch.full((b, n), 1e10, device=cuda)
    idx = torch.zeros((b, m), dtype=torch.int32, device=cuda)
    ops.farthest_point_sampling(b, n, m, pts, temp, idx)
    for f in range(b):
        i = idx[f].long()
        assert int(i[0]) == 0 and torch.unique(i).numel() == m
        s = pts[f][i]
        d2 = ((s.unsqueeze(1) - s.unsqueeze(0)) ** 2).sum(-1)
        d2 = d2.masked_fill(torch.triu(torch.ones((m, m), dtype=torch.bool, device=cuda)), float("inf"))
        gap = d2.min(1).values[1:]                                              # sample t against samples 0 .. t-1
        assert bool((gap[1:] <= gap[:-1] * (1 + 1e-5)).all())


def test_ball_query_full_size_members_inside_radius(cuda):
    from crb3d import ops
    rng = np.random.default_rng(6)
    nb, n, m, radius, nsample = 4, 16384, 2048, 0.8, 16
    xyz = rng.uniform([0, -40, -3], [70, 40, 1], (nb * n, 3)).astype(np.float32)
    new_xyz = np.concatenate([xyz[f * n:f * n + m] + np.float32(0.01) for f in range(nb)])
    cnt, ncnt = np.full(nb, n, np.int32), np.full(nb, m, np.int32)
    idx = torch.zeros((nb * m, nsample), dtype=torch.int32, device=cuda)
    X, Q = cu(xyz, cuda), cu(new_xyz, cuda)
    ops.ball_query(nb, nb * m, radius, nsample, Q, cu(ncnt, cuda), X, cu(cnt, cuda), idx)
    idx = idx.long().view(nb, m, nsample)
    assert int(idx.min()) >= 0 and int(idx.max()) < n                          # every query has itself (shifted by 1 cm) nearby
    for f in range(nb):
        src = X[f * n:(f + 1) * n]
        q = Q[f * m:(f + 1) * m]
        d2 = ((src[idx[f]] - q.unsqueeze(1)) ** 2).sum(-1)
        assert float(d2.max()) < radius * radius
        # slot 0 is the lowest-index source inside the ball; slots never go backwards before the padding repeats slot 0
        inside = ((src.unsqueeze(0) - q.unsqueeze(1)) ** 2).sum(-1) < radius * radius - 1e-4
        first = torch.argmax(inside.int(), 1)
        assert bool((idx[f][:, 0] <= first).all())
        step = idx[f][:, 1:] - idx[f][:, :-1]
        assert bool(((step > 0) | (idx[f][:, 1:] == idx[f][:, :1])).all())


def test_points_in_boxes_full_size_consistent(cuda):
    from crb3d import ops
    rng = np.random.default_rng(8)
    nb, nbox, npts = 8, 128, 20000
    boxes = np.stack([rand_boxes(rng, nbox, 30) for _ in range(nb)])
    boxes[..., 0] += 35.0
    pts = rng.uniform([0, -40, -3], [70, 40, 1], (nb, npts, 3)).astype(np.float32)
    pts[:, :nbox] = boxes[..., :3]                                             # box centres are inside their own box
    Bx, Pt = cu(boxes, cuda), cu(pts, cuda)
    out = ops.points_in_boxes(Bx, Pt).long()
    assert bool((out[:, :nbox] >= 0).all()) and int(out.max()) < nbox and int(out.min()) == -1
    # in the box frame: an assigned point is inside (1e-4 slack), a point more than 1e-4 outside every box is unassigned
    d = Pt.double().unsqueeze(2) - Bx[..., :3].double().unsqueeze(1)           # (nb, npts, nbox, 3)
    ang = -Bx[..., 6].double().unsqueeze(1)
    lx = d[..., 0] * torch.cos(ang) - d[..., 1] * torch.sin(ang)
    ly = d[..., 0] * torch.sin(ang) + d[..., 1] * torch.cos(ang)
    hx, hy, hz = (Bx[..., 3 + j].double().unsqueeze(1) / 2 for j in range(3))
    def inside(eps):
        return (lx.abs() < hx + eps) & (ly.abs() < hy + eps) & (d[..., 2].abs() <= hz + eps)
    loose, strict = inside(1e-4), inside(-1e-4)
    got = torch.gather(loose, 2, out.clamp(min=0).unsqueeze(-1)).squeeze(-1)
    assert bool((got | (out < 0)).all())
    assert bool(((out >= 0) | ~strict.any(2)).all())


def test_label_entropy_bounds_and_permutation(cuda):
    from crb3d import ops
    rng = np.random.default_rng(9)
    frames = [rng.integers(1, 4, rng.integers(0, 500)) for _ in range(4096)]
    off = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    lab = np.concatenate(frames)
    e1 = ops.label_entropy(cu(lab, cuda, torch.int32), cu(off, cuda), 3)
    shuffled = np.concatenate([f[rng.permutation(len(f))] for f in frames])
    e2 = ops.label_entropy(cu(shuffled, cuda, torch.int32), cu(off, cuda), 3)
    assert torch.equal(e1, e2)
    assert float(e1.min()) >= 0.0 and float(e1.max()) <= np.log(3) + 1e-6
