"""-m gpu parity: CRB stage-1 entropy, stage-2 distance matrix, stage-3 greedy KDE/KL - ours vs the reference's own
library calls (torch Categorical, sklearn euclidean_distances / KernelDensity, scipy entropy) restated in oracle/crb.py."""
import numpy as np
import pytest
import torch

from util import cu

pytestmark = pytest.mark.gpu


def test_label_entropy(cuda):
    from crb3d import ops
    from oracle import crb as oc
    rng = np.random.default_rng(1)
    frames = [rng.integers(1, 4, rng.integers(1, 80)) for _ in range(200)]
    frames += [np.zeros((0,), np.int64), np.array([2]), np.array([1] * 7 + [3] * 2), np.array([3] * 500), np.array([1, 2, 3])]
    off = np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)
    ent = ops.label_entropy(cu(np.concatenate(frames), cuda, torch.int32), cu(off, cuda), 3).cpu().numpy()
    ref = np.array([oc.label_entropy(f, 3) for f in frames], np.float32)
    assert np.allclose(ent, ref, rtol=1e-6, atol=1e-7)              # 1e-3 rel is the bar; fp32 rounding only
    assert ent[-5] == 0.0 and abs(ent[-3] - 0.8018185) < 1e-6
    # ranking identical to the reference's stable-sort-then-reverse (crb_sampling.py:119-121)
    ids = list(range(len(frames)))
    assert oc.stage1_shortlist(ids, list(ent), 50) == oc.stage1_shortlist(ids, list(ref), 50)


def test_pairwise_sqdist(cuda):
    from crb3d import ops
    from sklearn.metrics.pairwise import euclidean_distances
    rng = np.random.default_rng(2)
    X = rng.normal(size=(150, 4099)).astype(np.float32)
    D = ops.pairwise_sqdist(cu(X, cuda)).cpu().numpy()
    ref = euclidean_distances(X, X, squared=True)
    assert np.allclose(D, ref, rtol=1e-6, atol=1e-3)


def _pool(rng, n_frames, empty_class_rate=0.2):
    dens, labs = [], []
    for _ in range(n_frames):
        n = int(rng.integers(0, 40))
        l = rng.integers(1, 4, n)
        if rng.uniform() < empty_class_rate:
            l[l == 2] = 1
        d = np.where(l == 1, rng.gamma(2.0, 8.0, n), np.where(l == 2, rng.gamma(2.0, 30.0, n), rng.gamma(2.0, 15.0, n)))
        dens.append(d.astype(np.float32)); labs.append(l.astype(np.int64))
    return dens, labs


@pytest.mark.parametrize("seed,n_cand,n_sel", [(0, 24, 8), (1, 40, 12), (2, 12, 12)])
def test_kde_greedy_vs_sklearn(cuda, seed, n_cand, n_sel):
    from crb3d import crb_host, ops
    from oracle import crb as oc
    rng = np.random.default_rng(seed)
    pool_d, pool_l = _pool(rng, 300)
    x_axis, prior = oc.build_prior(np.concatenate(pool_d), np.concatenate(pool_l), 3)
    ax2, prior2 = crb_host.build_prior(torch.as_tensor(np.concatenate(pool_d)), torch.as_tensor(np.concatenate(pool_l)), 3)
    assert np.allclose(np.stack(x_axis), ax2.numpy(), rtol=0, atol=0)
    assert np.allclose(np.stack(prior), prior2.numpy(), rtol=0, atol=0, equal_nan=True)
    cand_d, cand_l = pool_d[:n_cand], pool_l[:n_cand]
    picked_o, score_o = oc.greedy_density_balance(list(cand_d), list(cand_l), x_axis, prior, 3, n_sel, bandwidth=5)
    off = np.cumsum([0] + [len(d) for d in cand_d]).astype(np.int32)
    prior_n = crb_host.normalise_prior(prior2)
    order, ps = ops.kde_greedy(cu(np.concatenate(cand_d), cuda), cu(np.concatenate(cand_l), cuda, torch.int32), cu(off, cuda),
                               3, ax2.to(cuda), prior_n.to(cuda), 5.0, n_sel)
    assert order.cpu().tolist() == picked_o                                   # identical greedy selection order
    got = ps.cpu().numpy()[1:]
    want = np.asarray(score_o[1:], np.float64)
    assert np.allclose(got, want, rtol=1e-6, atol=1e-9)                       # CRB scores (bar: 1e-3 rel)


def test_crb_query_end_to_end_second(cuda):
    """CRBSampling.query on a small synthetic pool with a SECOND detector: every stage is re-derived with the reference's
    own library calls (oracle/crb.py) from the same per-frame records / embeddings and must select the same frames."""
    from crb3d import crb_strategy, scorer, second, synth
    from oracle import crb as oc
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(cuda)
    frames = {"%06d" % i: synth.make_frame(100 + i)[::2] for i in range(14)}          # ~10k points per frame
    ps = scorer.PoolScorer(model, cuda, batch_size=4)
    first = ps.to_device(ps.stage_host(list(frames.values())[:4]))
    second.calibrate_head_bias(model, first[0], first[1], 4, target_fraction=0.004)
    cfg = {"ACTIVE_TRAIN": {"SELECT_NUMS": 3, "ACTIVE_CONFIG": {"K1": 3, "K2": 2, "BANDWIDTH": 7}}}   # correct spelling is ignored
    strat = crb_strategy.CRBSampling(model, None, frames, 0, "/tmp", cfg)
    assert strat.bandwidth == 5                                                     # the BANDWDITH quirk
    selected = strat.query(cur_epoch=0, scorer=ps)
    st = strat.last_stage
    recs = st["records"]
    assert len(selected) == 3 and len(set(selected)) == 3 and set(selected) <= set(frames)
    ids = list(frames.keys())
    # stage 1: entropy per frame == torch Categorical on the frame's labels; shortlist == stable sort, reversed
    ent_o = [oc.label_entropy(recs[i]["labels"], 3) for i in ids]
    assert np.allclose([recs[i]["entropy"] for i in ids], ent_o, rtol=1e-5, atol=1e-6)
    assert st["shortlist"] == oc.stage1_shortlist(ids, [recs[i]["entropy"] for i in ids], 9)
    # stage 2: k-means++ seeding on the embedding matrix == sklearn on the same matrix
    emb = st["embeddings"].cpu().numpy()
    assert emb.shape == (9, 18 * 512) and np.isfinite(emb).all() and np.abs(emb).max() > 0
    want = oc.kmeanspp_indices(emb, 6)
    assert st["prototypes"] == [st["shortlist"][i] for i in want]
    # stage 3: greedy KDE/KL on the records == the sklearn/scipy loop
    x_axis, prior = oc.build_prior(np.concatenate([recs[i]["density"] for i in ids]), np.concatenate([recs[i]["labels"] for i in ids]), 3)
    picked, _ = oc.greedy_density_balance([recs[f]["density"] for f in st["prototypes"]], [recs[f]["labels"] for f in st["prototypes"]],
                                          x_axis, prior, 3, 3, bandwidth=5)
    assert selected == [st["prototypes"][i] for i in picked]


def test_second_gradient_embedding_is_conv_cls_weight_grad(cuda):
    """The closed-form last-layer embedding (delta^T X) == autograd's conv_cls.weight.grad of the focal loss."""
    from crb3d import crb_strategy, scorer, second, synth
    torch.manual_seed(1)
    model = second.SECONDNet().eval().to_device(cuda)
    ps = scorer.PoolScorer(model, cuda, batch_size=1)
    pts = synth.make_frame(7)[::3]
    strat = crb_strategy.CRBSampling(model, None, {}, 0, "/tmp", {})
    emb = strat.second_gradient_embedding(ps, pts)
    with torch.no_grad():
        p = torch.from_numpy(pts).to(cuda)
        offs = torch.tensor([0, len(pts)], dtype=torch.int32, device=cuda)
        bd = model.forward_features(p, offs, 1)
    x = bd["spatial_features_2d"].detach()
    w = model.dense_head.conv_cls.weight.detach().clone().requires_grad_(True)
    logits = torch.nn.functional.conv2d(x, w, model.dense_head.conv_cls.bias).permute(0, 2, 3, 1).reshape(-1, 3)
    labels = torch.argmax(logits, -1)
    target = torch.zeros_like(logits)
    pos = labels > 0
    target[pos, labels[pos] - 1] = 1.0
    prob = torch.sigmoid(logits)
    aw = target * 0.25 + (1 - target) * 0.75
    pt = target * (1 - prob) + (1 - target) * prob
    bce = torch.clamp(logits, min=0) - logits * target + torch.log1p(torch.exp(-torch.abs(logits)))
    loss = (aw * pt ** 2 * bce).sum() / torch.clamp(pos.sum().float(), min=1.0)
    loss.backward()
    ref = w.grad.reshape(-1)
    assert float((emb - ref).abs().max()) <= 2e-3 * float(ref.abs().max())


def test_coreset_furthest_first_and_badge_kmeanspp(cuda):
    """crb3d.strategies against a numpy restatement of coreset_sampling.py:31-52 (mean-initialised greedy k-centre) and
    sklearn.cluster.kmeans_plusplus (badge_sampling.py:190-196)."""
    from sklearn.cluster import kmeans_plusplus
    from crb3d import strategies
    torch.backends.cuda.matmul.allow_tf32 = False
    rng = np.random.default_rng(5)
    X = rng.normal(size=(400, 64)).astype(np.float32) * rng.uniform(0.5, 3.0, size=(400, 1)).astype(np.float32)
    Xs = rng.normal(size=(50, 64)).astype(np.float32)

    def sq(a, b):
        d = (a.astype(np.float64) ** 2).sum(1)[:, None] + (b.astype(np.float64) ** 2).sum(1)[None, :] - 2.0 * a.astype(np.float64) @ b.astype(np.float64).T
        return np.clip(d, 0, None)
    min_dist = sq(X, Xs).mean(1)
    ref = []
    for i in range(30):
        j = int(np.argmax(min_dist))
        ref.append(j)
        min_dist = np.minimum(min_dist, sq(X, X[j:j + 1])[:, 0])
    got = strategies.furthest_first(torch.from_numpy(X).to(cuda), torch.from_numpy(Xs).to(cuda), 30).cpu().numpy()
    assert np.array_equal(got, np.asarray(ref))
    _, idx = kmeans_plusplus(X, 25, random_state=0)
    assert np.array_equal(strategies.kmeans_pp_select(torch.from_numpy(X).to(cuda), 25), idx)
