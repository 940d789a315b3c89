"""-m gpu parity: voxelize+MeanVFE, rulebook, sparse conv fwd/bwd, dense - CUDA (through the C ABI) vs the oracle."""
import numpy as np
import pytest
import torch

from util import cu

pytestmark = pytest.mark.gpu

K_RANGE = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
K_VOX = [0.05, 0.05, 0.1]


def _stack(frames):
    offs = np.zeros(len(frames) + 1, np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    return np.concatenate(frames) if sum(len(f) for f in frames) else np.zeros((0, frames[0].shape[1]), np.float32), offs


def _check_voxelize(frames, pc_range, vsize, max_pts, max_voxels, dev):
    from crb3d import ops
    from oracle import voxel
    pts, offs = _stack(frames)
    res = ops.voxelize(cu(pts, dev), cu(offs, dev), len(frames), pc_range, vsize, max_pts, max_voxels, want_voxels=True)
    mean_o, coords_o, num_o, voff_o = voxel.voxelize_batch(frames, pc_range, vsize, max_pts, max_voxels)
    vox_o = np.concatenate([voxel.point_to_voxel(f, pc_range, vsize, max_pts, max_voxels)[0] for f in frames])
    assert np.array_equal(res["frame_voxel_offsets"].cpu().numpy(), voff_o)
    assert np.array_equal(res["coords"].cpu().numpy(), coords_o)          # bit-exact indices + first-seen order
    assert np.array_equal(res["num_points"].cpu().numpy(), num_o)
    assert np.array_equal(res["voxels"].cpu().numpy(), vox_o)              # same points in the same slots
    assert np.array_equal(res["mean"].cpu().numpy(), mean_o)               # same summation order -> bit-exact
    return res


def test_voxelize_kitti_frames(cuda):
    from crb3d import synth
    frames = [synth.make_frame(i) for i in range(3)]
    res = _check_voxelize(frames, K_RANGE, K_VOX, 5, 40000, cuda)
    assert 14000 < res["coords"].shape[0] / 3 < 19000


def test_voxelize_edge_cases(cuda):
    rng = np.random.default_rng(7)
    f0 = np.concatenate([rng.uniform([-5, -45, -4], [75, 45, 2], (3000, 3)), rng.uniform(0, 1, (3000, 1))], 1).astype(np.float32)
    f0[500:1500] = f0[:1000] + np.float32(1e-4)        # many multi-point voxels (exercise the 5-point cap)
    f0[1500:1600] = f0[0]                               # 101 points in one voxel
    f1 = np.zeros((0, 4), np.float32)                   # empty frame
    f2 = f0[::-1].copy()
    f2[10, 0] = np.float32(70.4)                        # exactly on the upper bound -> rejected
    f2[11, :3] = np.float32([0.0, -40.0, -3.0])         # exactly on the lower bound -> voxel (0,0,0)
    _check_voxelize([f0, f1, f2], K_RANGE, K_VOX, 5, 40000, cuda)
    _check_voxelize([f0, f1, f2], K_RANGE, K_VOX, 5, 700, cuda)    # max_voxels truncation (first-seen voxels win)
    _check_voxelize([f0], K_RANGE, [0.4, 0.4, 0.4], 3, 100000, cuda)
    _check_voxelize([f1], K_RANGE, K_VOX, 5, 100, cuda)


def _rand_coords(rng, B, shape, n):
    c = np.stack([rng.integers(0, B, n), rng.integers(0, shape[0], n), rng.integers(0, shape[1], n), rng.integers(0, shape[2], n)], 1)
    c = np.unique(c, axis=0).astype(np.int32)
    return c[rng.permutation(len(c))]


@pytest.mark.parametrize("shape,n", [([41, 160, 140], 30000), ([5, 7, 3], 60), ([9, 24, 20], 1)])
def test_subm_rulebook(cuda, shape, n):
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(n)
    coords = _rand_coords(rng, 3, shape, n)
    nbr = ops.subm_rulebook(cu(coords, cuda), shape, (3, 3, 3))
    ref = spconv_ref.subm_rulebook(coords, shape, (3, 3, 3))
    assert np.array_equal(nbr.cpu().numpy(), ref)
    pairs, num = ops.compact_pairs(nbr)
    pref, nref = spconv_ref.pairs_from_table(ref)
    assert np.array_equal(num.cpu().numpy(), nref) and np.array_equal(pairs.cpu().numpy(), pref)


@pytest.mark.parametrize("k,s,p", [((3, 3, 3), (2, 2, 2), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (0, 1, 1)),
                                   ((3, 1, 1), (2, 1, 1), (0, 0, 0)), ((2, 2, 2), (2, 2, 2), (0, 0, 0))])
def test_sparse_rulebook(cuda, k, s, p):
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(11)
    shape = [21, 80, 72]
    coords = _rand_coords(rng, 2, shape, 20000)
    oc, oshape, nbr, nbr_t = ops.sparse_rulebook(cu(coords, cuda), 2, shape, k, s, p)
    roc, roshape, rnbr, rnbr_t = spconv_ref.sparse_rulebook(coords, 2, shape, k, s, p)
    assert oshape == roshape
    assert np.array_equal(oc.cpu().numpy(), roc)       # ascending (b,z,y,x) key order, bit-exact
    assert np.array_equal(nbr.cpu().numpy(), rnbr)
    assert np.array_equal(nbr_t.cpu().numpy(), rnbr_t)


@pytest.mark.parametrize("k,s,p,ks", [((3, 3, 3), (2, 2, 2), (1, 1, 1), (3, 3, 3)), ((3, 3, 3), (2, 2, 2), (0, 1, 1), (3, 3, 3)),
                                      ((3, 1, 1), (2, 1, 1), (0, 0, 0), (3, 3, 3)), ((3, 3, 3), (2, 2, 2), (1, 1, 1), (1, 3, 5))])
def test_subm_rulebook_through_cellmap(cuda, k, s, p, ks):
    """SubM table of a strided level read off the producing rulebook's bitmap ranks == the hash-table route == the oracle;
    also with device-side counts on capacity-sized buffers, and with a capacity that truncates the level."""
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(sum(k) + sum(p) + sum(ks))
    shape = [21, 80, 72]
    coords = _rand_coords(rng, 3, shape, 25000)
    ct = cu(coords, cuda)
    oc, oshape, nbr, nbr_t, cm = ops.sparse_rulebook(ct, 3, shape, k, s, p, want_cellmap=True)
    via_map = ops.subm_rulebook(oc, oshape, ks, cellmap=cm)
    via_hash = ops.subm_rulebook(oc, oshape, ks)
    assert torch.equal(via_map, via_hash)
    assert np.array_equal(via_map.cpu().numpy(), spconv_ref.subm_rulebook(oc.cpu().numpy(), oshape, ks))
    assert not cm.matches(ct, shape)                       # a map only answers for the rows it ranked
    # static route: padded input, device-side counts, capacity above / below the true output count
    n_out = oc.shape[0]
    pad = torch.cat([ct, torch.full((777, 4), 3, dtype=torch.int32, device=cuda)])
    n_in_dev = torch.tensor([ct.shape[0]], dtype=torch.int32, device=cuda)
    for cap in (n_out + 500, n_out - 1000):
        soc, soshape, snbr, n_dev, scm = ops.sparse_rulebook_static(pad, n_in_dev, 3, shape, k, s, p, cap, want_cellmap=True)
        assert int(n_dev.item()) == n_out and soshape == oshape
        m = min(cap, n_out)
        assert torch.equal(soc[:m], oc[:m]) and torch.equal(snbr[:, :m], nbr[:, :m])
        t = ops.subm_rulebook(soc, soshape, ks, n_dev=n_dev, cellmap=scm)
        h = ops.subm_rulebook(soc, soshape, ks, n_dev=n_dev)
        if cap >= n_out:
            assert torch.equal(t[:, :m], via_map)
        else:                                               # truncated level: neighbours beyond the capacity read as absent
            expect = via_map[:, :m].clone()
            expect[expect >= m] = -1
            assert torch.equal(t[:, :m], expect)
        assert torch.equal(t[:, :m], h[:, :m]) or cap < n_out


def test_sparse_rulebook_empty(cuda):
    from crb3d import ops
    oc, oshape, nbr, nbr_t = ops.sparse_rulebook(torch.zeros((0, 4), dtype=torch.int32, device=cuda), 1, [9, 8, 8],
                                                 (3, 3, 3), (2, 2, 2), (1, 1, 1))
    assert oc.shape[0] == 0 and nbr.shape == (27, 0)


@pytest.mark.parametrize("cin,cout", [(4, 16), (16, 16), (16, 32), (32, 64), (64, 64), (64, 128), (5, 16)])
def test_spconv_forward_backward(cuda, cin, cout):
    """fp32 SIMT path vs the fp64 gather/mm/scatter oracle; tolerance 1e-5 relative to the output scale (fp32 rounding)."""
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(cin * 131 + cout)
    shape = [11, 40, 36]
    coords = _rand_coords(rng, 2, shape, 6000)
    n = len(coords)
    feat = rng.normal(size=(n, cin)).astype(np.float32)
    w = (rng.normal(size=(cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)
    for kind in ("subm", "sparse"):
        if kind == "subm":
            nbr_ref = spconv_ref.subm_rulebook(coords, shape, (3, 3, 3))
            nbr = ops.subm_rulebook(cu(coords, cuda), shape, (3, 3, 3))
            nbr_t = nbr.flip(0).contiguous()           # SubM: transpose table = offsets reversed
        else:
            _, _, nbr_ref, nbr_t_ref = spconv_ref.sparse_rulebook(coords, 2, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
            _, _, nbr, nbr_t = ops.sparse_rulebook(cu(coords, cuda), 2, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
        out = ops.spconv_forward(cu(feat, cuda), nbr, cu(w, cuda))
        ref = spconv_ref.conv_forward(feat, nbr_ref, w, dtype=torch.float64)
        scale = float(ref.abs().max())
        assert float((out.cpu().double() - ref).abs().max()) <= 1e-5 * scale
        dout = rng.normal(size=tuple(ref.shape)).astype(np.float32)
        dx_ref, dw_ref = spconv_ref.conv_backward(feat, nbr_ref, w, dout, dtype=torch.float64)
        dx = ops.spconv_forward(cu(dout, cuda), nbr_t, cu(w, cuda), transpose=True)
        if kind == "subm":  # the autograd path: forward table + reversed weight slices (no flipped copy of the table)
            kmap = torch.arange(26, -1, -1, dtype=torch.int32, device=cuda)
            dx2 = ops.spconv_forward(cu(dout, cuda), nbr, cu(w, cuda), transpose=True, kmap=kmap)
            assert torch.allclose(dx, dx2, rtol=1e-5, atol=1e-5)   # same pairs, offsets summed in reverse order
        dw = ops.spconv_wgrad(cu(feat, cuda), cu(dout, cuda), nbr, w.shape)
        assert float((dx.cpu().double() - dx_ref).abs().max()) <= 1e-5 * float(dx_ref.abs().max())
        assert float((dw.cpu().double() - dw_ref).abs().max()) <= 2e-5 * float(dw_ref.abs().max())


def test_spconv_fused_bn_relu(cuda):
    from crb3d import ops
    rng = np.random.default_rng(5)
    coords = _rand_coords(rng, 1, [9, 20, 20], 1500)
    feat = cu(rng.normal(size=(len(coords), 16)).astype(np.float32), cuda)
    w = cu((rng.normal(size=(32, 3, 3, 3, 16)) * 0.05).astype(np.float32), cuda)
    scale = cu(rng.uniform(0.5, 1.5, 32).astype(np.float32), cuda)
    shift = cu(rng.normal(size=32).astype(np.float32), cuda)
    nbr = ops.subm_rulebook(cu(coords, cuda), [9, 20, 20], (3, 3, 3))
    plain = ops.spconv_forward(feat, nbr, w)
    fused = ops.spconv_forward(feat, nbr, w, scale=scale, shift=shift, relu=True)
    assert torch.allclose(fused, torch.relu(plain * scale + shift), rtol=1e-6, atol=1e-6)


def test_dense_roundtrip(cuda):
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(9)
    shape = [2, 25, 22]
    coords = _rand_coords(rng, 3, shape, 900)
    feat = rng.normal(size=(len(coords), 128)).astype(np.float32)
    d = ops.sparse_to_dense(cu(feat, cuda), cu(coords, cuda), 3, shape)
    ref = spconv_ref.dense(feat, coords, 3, shape)
    assert torch.equal(d.cpu(), ref)
    cl = ops.sparse_to_dense(cu(feat, cuda), cu(coords, cuda), 3, shape, channels_last_bev=True)
    assert torch.equal(cl.permute(0, 3, 1, 2).cpu(), ref.view(3, 128 * 2, 25, 22))   # height_compression.py:21-23
    back = ops.dense_to_sparse(d, cu(coords, cuda), 128, shape)
    assert torch.equal(back.cpu(), torch.as_tensor(feat))


@pytest.mark.parametrize("cin,cout", [(16, 16), (16, 32), (32, 32), (32, 64), (64, 64), (64, 128)])
def test_spconv_tcgen05_tf32(cuda, cin, cout):
    """tcgen05 (TF32 in, fp32 accumulate) kernel vs the fp64 oracle. TF32 keeps 10 mantissa bits: the bound is
    ~2^-11 * sqrt(2) relative per product, averaged over the sum -> 1e-3 of the output scale (north_star tolerance)."""
    from crb3d import ops
    from oracle import spconv_ref
    rng = np.random.default_rng(cin * 7 + cout)
    shape = [11, 64, 60]
    coords = _rand_coords(rng, 2, shape, 20000)
    n = len(coords)
    feat = rng.normal(size=(n, cin)).astype(np.float32)
    w = (rng.normal(size=(cout, 3, 3, 3, cin)) / np.sqrt(27 * cin)).astype(np.float32)
    scale = rng.uniform(0.5, 1.5, cout).astype(np.float32)
    shift = rng.normal(size=cout).astype(np.float32)
    for kind in ("subm", "sparse", "k311"):
        if kind == "subm":
            nbr_ref = spconv_ref.subm_rulebook(coords, shape, (3, 3, 3))
            nbr = ops.subm_rulebook(cu(coords, cuda), shape, (3, 3, 3))
            ww = w
        elif kind == "sparse":
            _, _, nbr_ref, _ = spconv_ref.sparse_rulebook(coords, 2, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
            _, _, nbr, nbr_t = ops.sparse_rulebook(cu(coords, cuda), 2, shape, (3, 3, 3), (2, 2, 2), (1, 1, 1))
            ww = w
        else:
            _, _, nbr_ref, _ = spconv_ref.sparse_rulebook(coords, 2, shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
            _, _, nbr, _ = ops.sparse_rulebook(cu(coords, cuda), 2, shape, (3, 1, 1), (2, 1, 1), (0, 0, 0))
            ww = w[:, :, :1, :1, :].copy()
        ref = spconv_ref.conv_forward(feat, nbr_ref, ww, dtype=torch.float64)
        out = ops.spconv_forward(cu(feat, cuda), nbr, cu(ww, cuda), tf32=True)
        exact = ops.spconv_forward(cu(feat, cuda), nbr, cu(ww, cuda), tf32=False)
        s = float(ref.abs().max())
        err = float((out.cpu().double() - ref).abs().max())
        assert err <= 2e-3 * s, (kind, err / s)
        assert float((out - exact).abs().max()) > 0 or cin < 8          # it really is a different (TF32) datapath
        rel_rms = float(((out.cpu().double() - ref) ** 2).mean().sqrt() / (ref ** 2).mean().sqrt())
        assert rel_rms <= 1e-3, (kind, rel_rms)
        fused = ops.spconv_forward(cu(feat, cuda), nbr, cu(ww, cuda), scale=cu(scale, cuda), shift=cu(shift, cuda), relu=True, tf32=True)
        assert torch.allclose(fused, torch.relu(out * cu(scale, cuda) + cu(shift, cuda)), rtol=1e-5, atol=1e-5)
        if kind == "sparse":   # input gradient through the transposed table with the transposed weight
            dout = rng.normal(size=tuple(ref.shape)).astype(np.float32)
            dx_ref, _ = spconv_ref.conv_backward(feat, nbr_ref, ww, dout, dtype=torch.float64)
            dx = ops.spconv_forward(cu(dout, cuda), nbr_t, cu(ww, cuda), transpose=True, tf32=True)
            assert float((dx.cpu().double() - dx_ref).abs().max()) <= 2e-3 * float(dx_ref.abs().max())


def test_mask_collate_points_matches_numpy(cuda):
    """mask_points_by_range (common_utils.py:60-63: x and y only, closed interval) + collate_batch's batch column
    (dataset.py:173-178) for a whole batch on the device: same rows, same order as the numpy code of the reference."""
    from crb3d import ops
    rng = np.random.default_rng(4)
    pc_range = [0.0, -40.0, -3.0, 70.4, 40.0, 1.0]
    frames = []
    for n in (5000, 1, 0, 7321):
        f = rng.uniform([-10, -50, -5, 0], [80, 50, 3, 1], (n, 4)).astype(np.float32)
        if n > 10:
            f[0, 0], f[1, 0], f[2, 1], f[3, 1] = 0.0, 70.4, -40.0, 40.0      # on the boundary: kept (closed interval)
            f[4, 0] = np.nan
            f[5, 2] = 100.0                                                    # z is not tested
        frames.append(f)
    offs = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    pts = np.concatenate(frames)
    out, out_off = ops.mask_collate_points(torch.from_numpy(pts).to(cuda), torch.from_numpy(offs).to(cuda), pc_range)
    ref, ref_off = [], [0]
    for i, f in enumerate(frames):
        m = (f[:, 0] >= pc_range[0]) & (f[:, 0] <= pc_range[3]) & (f[:, 1] >= pc_range[1]) & (f[:, 1] <= pc_range[4])
        ref.append(np.pad(f[m], ((0, 0), (1, 0)), mode="constant", constant_values=i))
        ref_off.append(ref_off[-1] + int(m.sum()))
    ref = np.concatenate(ref)
    assert np.array_equal(out_off.cpu().numpy(), np.asarray(ref_off, np.int32))
    assert np.array_equal(out.cpu().numpy(), ref, equal_nan=True)


def test_raw_pointer_wrappers_reject_wrong_dtype_and_layout(cuda):
    """CHECK_INPUT of the reference's pybind layer (CUDA + contiguous) plus the dtype: an int64 count vector or a strided view
    must raise instead of producing garbage indices."""
    from crb3d import ops
    xyz = torch.rand(100, 3, device=cuda)
    idx = torch.zeros((10, 4), dtype=torch.int32, device=cuda)
    cnt64 = torch.tensor([100], dtype=torch.int64, device=cuda)
    cnt = torch.tensor([100], dtype=torch.int32, device=cuda)
    ncnt = torch.tensor([10], dtype=torch.int32, device=cuda)
    with pytest.raises(TypeError):
        ops.ball_query(1, 10, 0.5, 4, xyz[:10].contiguous(), ncnt, xyz, cnt64, idx)
    with pytest.raises(ValueError):
        ops.ball_query(1, 10, 0.5, 4, xyz[::10], ncnt, xyz, cnt, idx)
    with pytest.raises(RuntimeError):
        ops.ball_query(1, 10, 0.5, 4, xyz[:10].cpu(), ncnt, xyz, cnt, idx)
    # a batch size that does not match the offsets vector would read past it on the device
    with pytest.raises(ValueError):
        ops.voxelize(torch.rand(100, 4, device=cuda), torch.tensor([0, 50, 100], dtype=torch.int32, device=cuda), 4,
                     [0, -40, -3, 70.4, 40, 1], [0.05, 0.05, 0.1], 5, 1000)
