"""CPU tests: the oracle's C restatement against golden vectors produced by the REFERENCE's own CUDA kernels
(tests/golden/ref_kernels_*.npz, generated on a B200 by tests/golden/make_golden_gpu.py from oracle/_ref), plus the
oracle's internal cross-checks (python vs C voxelizer, gather/mm/scatter vs dense conv3d)."""
import os

import numpy as np
import pytest
import torch

from util import GOLDEN


def _load(name):
    path = os.path.join(GOLDEN, name)
    if not os.path.exists(path):
        pytest.skip("golden file %s missing" % name)
    return np.load(path)


def test_iou3d_golden():
    from oracle import boxes as ob
    g = _load("ref_kernels_iou3d.npz")
    # libm vs CUDA sincos/atan2 differ by ulps: values to 1e-4 abs (areas O(10)), NMS keep lists exact (margin checked)
    assert np.abs(ob.boxes_overlap_bev(g["a"], g["b"]) - g["overlap"]).max() < 1e-3
    assert np.abs(ob.boxes_iou_bev(g["a"], g["b"]) - g["iou"]).max() < 1e-4
    keep, iou = ob.nms_sorted(g["nms_boxes"], float(g["thr_rot"]), rotated=True, return_iou=True)
    assert np.abs(iou[np.triu_indices(len(iou), 1)] - g["thr_rot"]).min() > 1e-6
    assert np.array_equal(keep, g["keep_rot"])
    keep_n = ob.nms_sorted(g["nms_boxes"], float(g["thr_normal"]), rotated=False)
    assert np.array_equal(keep_n, g["keep_normal"])


def test_roiaware_golden():
    from oracle import boxes as ob
    g = _load("ref_kernels_roiaware.npz")
    for b in range(g["boxes"].shape[0]):
        mine = ob.points_in_boxes(g["boxes"][b], g["pts"][b])
        assert (mine != g["pib"][b]).mean() < 5e-4          # only points that sit on a face within an ulp may differ
    pooled, argmax, pidx = ob.roiaware_pool3d(g["rois"], g["pts"][0], g["feat"], 4, 10, "max")
    assert (pidx != g["pidx"]).mean() < 1e-3
    if np.array_equal(pidx, g["pidx"]):
        assert np.array_equal(argmax, g["argmax"]) and np.allclose(pooled, g["pooled_max"])
        pooled_a, _, _ = ob.roiaware_pool3d(g["rois"], g["pts"][0], g["feat"], 4, 10, "avg")
        assert np.allclose(pooled_a, g["pooled_avg"], rtol=1e-6, atol=1e-6)


def test_pointnet2_golden():
    """No transcendental is involved, and the oracle spells out the reference SASS's FMA contraction: bit-exact."""
    from oracle import pointnet2 as op
    g = _load("ref_kernels_pointnet2.npz")
    idx = op.ball_query(float(g["bq_radius"]), 16, g["xyz"], g["xyz_cnt"], g["new_xyz"], g["new_cnt"])
    assert np.array_equal(idx, g["bq_idx"])
    for b in range(2):
        fi, _ = op.farthest_point_sampling(g["fps_pts"][b], g["fps_idx"].shape[1])
        assert np.array_equal(fi, g["fps_idx"][b])
    fi2, _ = op.farthest_point_sampling(g["fps_pts2"][0], g["fps_idx2"].shape[1])     # n=700 -> reference block 512
    assert np.array_equal(fi2, g["fps_idx2"][0])
    d2, nidx = op.three_nn(g["unknown"], g["uc"], g["known"], g["kc"])
    assert np.array_equal(nidx, g["nn_idx"]) and np.array_equal(d2, g["nn_d2"])


def test_extra_ops_golden():
    """voxel query / roipoint pooling / batch-layout ball query + 3-NN of the oracle vs the reference kernels' outputs."""
    from oracle import pointnet2 as op
    g = _load("ref_kernels_extra.npz")
    idx = op.voxel_query(tuple(int(v) for v in g["vq_range"]), float(g["vq_radius"]), 16, g["vq_xyz"], g["vq_new_xyz"], g["vq_coords"],
                         g["vq_table"])
    assert np.array_equal(idx, g["vq_idx"])
    assert (idx[:, 0] == -1).any() and (idx[:, 0] >= 0).any()
    pooled, empty = op.roipoint_pool3d(g["rp_pts"], g["rp_boxes"], g["rp_feat"], g["rp_pooled"].shape[1])
    assert np.array_equal(empty, g["rp_empty"]) and empty[0] == 1 and empty[1] == 0
    # host libm sinf/cosf vs CUDA: a point exactly on a face may flip -> at most one box may differ
    assert (np.abs(pooled - g["rp_pooled"]).reshape(len(empty), -1).max(1) > 0).sum() <= 1
    n, m = g["bq_xyz"].shape[1], g["bq_new"].shape[1]
    for b in range(2):
        o = op.ball_query(float(g["bq_radius"]), 16, g["bq_xyz"][b], np.array([n], np.int32), g["bq_new"][b], np.array([m], np.int32))
        o[o[:, 0] == -1] = 0                       # the batch layout keeps the caller's zeros for an empty ball
        assert np.array_equal(o, g["bq_idx"][b])
        d2, nidx = op.three_nn(g["bq_new"][b], np.array([m], np.int32), g["bq_xyz"][b], np.array([n], np.int32))
        assert np.array_equal(nidx, g["nn_idx"][b]) and np.array_equal(d2, g["nn_d2"][b])


def test_fps_tie_rule_matches_reference_tree():
    """Exact duplicates: the reference's tree reduction ranks tied threads by bit-reversed thread id."""
    from oracle import pointnet2 as op
    rng = np.random.default_rng(0)
    pts = rng.uniform(-5, 5, (1024, 3)).astype(np.float32)
    pts[46] = pts[56] = np.float32([100, 100, 100])      # farthest point, duplicated in threads 46 and 56
    idx, _ = op.farthest_point_sampling(pts, 2)
    assert idx[1] == 56                                    # 46^56 = 0b010110 -> lowest differing bit 1: 56 has a 0 there


def test_voxelizer_c_vs_python_and_edges():
    from oracle import voxel
    rng = np.random.default_rng(1)
    R, V = [0, -40, -3, 70.4, 40, 1], [0.05, 0.05, 0.1]
    pts = np.concatenate([rng.uniform([-2, -42, -3.5], [72, 42, 1.5], (4000, 3)), rng.uniform(0, 1, (4000, 1))], 1).astype(np.float32)
    pts[1000:1400] = pts[:400] + np.float32(1e-4)
    pts[5] = [70.4, 0, 0, 0.5]          # on the upper bound -> dropped
    pts[6] = [0, -40, -3, 0.5]          # on the lower bound -> voxel 0,0,0
    for mv in (100000, 300):
        a = voxel.point_to_voxel(pts, R, V, 5, mv)
        b = voxel.point_to_voxel_py(pts, R, V, 5, mv)
        assert all(np.array_equal(x, y) for x, y in zip(a, b))
        assert a[2].max() <= 5 and len(a[1]) <= mv
    assert len(np.unique(a[1], axis=0)) == len(a[1])
    v, c, n = voxel.point_to_voxel(np.zeros((0, 4), np.float32), R, V, 5, 10)
    assert v.shape == (0, 5, 4) and c.shape == (0, 3)
    m = voxel.mean_vfe(a[0], a[2])
    assert np.allclose(m, a[0].sum(1) / np.maximum(a[2], 1)[:, None], rtol=1e-6)


@pytest.mark.parametrize("k,s,p", [((3, 3, 3), (1, 1, 1), (1, 1, 1)), ((3, 3, 3), (2, 2, 2), (1, 1, 1)),
                                   ((3, 3, 3), (2, 2, 2), (0, 1, 1)), ((3, 1, 1), (2, 1, 1), (0, 0, 0))])
def test_sparse_conv_oracle_vs_dense_conv3d(k, s, p):
    """Pins rulebook + conv semantics without spconv (PARITY UNPINNED at the library boundary): scatter to a dense grid,
    torch.nn.functional.conv3d, read back at the active sites (SURVEY.md 2.4 last bullet)."""
    from oracle import spconv_ref
    rng = np.random.default_rng(2)
    B, shape = 2, [9, 20, 18]
    coords = np.unique(np.stack([rng.integers(0, B, 500), rng.integers(0, 9, 500), rng.integers(0, 20, 500),
                                 rng.integers(0, 18, 500)], 1), axis=0).astype(np.int32)
    coords = coords[rng.permutation(len(coords))]
    feat = rng.normal(size=(len(coords), 6)).astype(np.float32)
    w = rng.normal(size=(10, *k, 6)).astype(np.float32)
    subm = s == (1, 1, 1)
    if subm:
        nbr = spconv_ref.subm_rulebook(coords, shape, k)
        oc = coords
    else:
        oc, osh, nbr, nbr_t = spconv_ref.sparse_rulebook(coords, B, shape, k, s, p)
        assert np.all(np.diff(spconv_ref._key(oc, osh)) > 0)                     # ascending unique keys
        for kk in range(nbr.shape[0]):                                           # transpose table consistency
            o = np.nonzero(nbr[kk] >= 0)[0]
            assert np.array_equal(nbr_t[kk, nbr[kk, o]], o)
    out = spconv_ref.conv_forward(feat, nbr, w, dtype=torch.float64)
    dc, dout = spconv_ref.dense_conv_reference(feat, coords, B, shape, w, s, p, subm=subm)
    assert np.array_equal(dc, oc)
    assert float((out - dout).abs().max()) < 1e-10
    # backward against autograd through the dense conv
    g = torch.from_numpy(rng.normal(size=tuple(out.shape)))
    dx, dw = spconv_ref.conv_backward(feat, nbr, w, g, dtype=torch.float64)
    x = spconv_ref.dense(torch.from_numpy(feat).double(), coords, B, shape).requires_grad_(True)
    wt = torch.from_numpy(w).double().requires_grad_(True)
    y = torch.nn.functional.conv3d(x, wt.permute(0, 4, 1, 2, 3), stride=s, padding=p)
    c = torch.from_numpy(np.asarray(oc)).long()
    (y[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] * g).sum().backward()
    ci = torch.from_numpy(coords).long()
    assert float((x.grad[ci[:, 0], :, ci[:, 1], ci[:, 2], ci[:, 3]] - dx).abs().max()) < 1e-10
    assert float((wt.grad - dw).abs().max()) < 1e-9


def test_crb_oracle_closed_forms():
    """KDE log-density and KL closed forms used by the CUDA kernels == the sklearn / scipy calls of the reference."""
    import scipy.stats
    from sklearn.neighbors import KernelDensity
    rng = np.random.default_rng(3)
    d = rng.gamma(2.0, 10.0, 37).astype(np.float32)
    x = np.linspace(-50, 160, 400)
    lp = KernelDensity(kernel="gaussian", bandwidth=5).fit(d[:, None]).score_samples(x[:, None])
    u = (x[:, None] - d[None, :].astype(np.float64)) / 5.0
    e = -0.5 * u * u
    m = e.max(1)
    mine = m + np.log(np.exp(e - m[:, None]).sum(1)) - np.log(len(d) * 5.0 * np.sqrt(2 * np.pi))
    assert np.abs(mine - lp).max() < 1e-10
    pk = scipy.stats.uniform.pdf(x, 3, 60)
    q = np.exp(mine)
    kl = np.sum(np.where(pk > 0, (pk / pk.sum()) * np.log((pk / pk.sum()) / (q / q.sum())), 0.0))
    assert abs(kl - scipy.stats.entropy(pk, np.exp(lp))) < 1e-10
    from oracle import crb as oc
    assert abs(oc.label_entropy([1] * 7 + [3] * 2, 3) - 0.8018185) < 1e-6       # pseudo-count quirk
    assert oc.label_entropy([], 3) == 0.0
