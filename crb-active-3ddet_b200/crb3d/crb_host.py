"""Host-side logic of CRB acquisition that surrounds the CUDA scoring kernels (mirrors
pcdet/query_strategies/crb_sampling.py of the reference; every quirk that defines results is kept - SURVEY.md 2.5).

Nothing here touches oracle/ or sklearn/scipy at run time: the closed forms are restated and the tests check them
against the library calls the reference makes.
"""
import math

import numpy as np
import torch


def shortlist_by_entropy(frame_ids, entropies, k):
    """crb_sampling.py:119-121: dict(sorted(items, key=value)) is a stable ascending sort; the key list is reversed and
    cut to k => descending entropy, ties in REVERSE insertion order."""
    ent = np.asarray([float(e) for e in entropies], dtype=np.float64)
    order = np.argsort(ent, kind="stable")[::-1][:k]
    return [frame_ids[i] for i in order]


def build_prior(density_all, label_all, num_class, alpha=0.95):
    """crb_sampling.py:252-260. Returns (x_axis (C,400) f64, uniform pdf (C,400) f64) as CPU tensors.
    Raises IndexError like the reference when a class never occurs in the pool."""
    density_all = density_all.detach().float().cpu()
    label_all = label_all.detach().long().cpu()
    unique_labels, label_counts = torch.unique(label_all, return_counts=True)
    sorted_density = [torch.sort(density_all[label_all == u])[0] for u in unique_labels]
    if len(unique_labels) < num_class:
        raise IndexError("class list shorter than num_class (reference crb_sampling.py:259 indexes range(num_class))")
    gmax, ghigh, glow = [], [], []
    for u in range(len(unique_labels)):
        sd = sorted_density[u]
        k = int(alpha * label_counts[u])           # python float * int64 0-d tensor -> float32 tensor -> int()
        gmax.append(int(sd[-1]))
        ghigh.append(int(sd[k]))
        glow.append(int(sd[-k]))
    axis = np.stack([np.linspace(-50, int(gmax[i]) + 50, 400) for i in range(num_class)])
    prior = np.stack([uniform_pdf(axis[i], glow[i], ghigh[i] - glow[i]) for i in range(num_class)])
    return torch.from_numpy(axis), torch.from_numpy(prior)


def uniform_pdf(x, loc, scale):
    """scipy.stats.uniform.pdf(x, loc, scale): 1/scale on the closed support [loc, loc+scale], 0 outside, nan if scale<=0."""
    x = np.asarray(x, dtype=np.float64)
    if not scale > 0:
        return np.full_like(x, np.nan)
    y = (x - loc) / scale
    return np.where((y >= 0) & (y <= 1), 1.0 / scale, 0.0)


def normalise_prior(prior):
    """scipy.stats.entropy normalises pk to sum 1 (row-wise here)."""
    prior = prior.double()
    return prior / prior.sum(dim=1, keepdim=True)


def kmeans_plusplus_indices(sqdist, n_clusters, seed=0, dtype=np.float32):
    """sklearn.cluster.kmeans_plusplus(X, n_clusters, random_state=seed)[1] restated on a precomputed squared-distance
    matrix (from crb3d.ops.pairwise_sqdist). sklearn works in X's dtype (float32 for the gradient embeddings,
    crb_sampling.py:226), hence `dtype`; the RandomState call sequence (choice, then uniform per centre) is identical."""
    D = np.asarray(sqdist).astype(dtype)
    n = D.shape[0]
    rs = np.random.RandomState(seed)
    n_local_trials = 2 + int(np.log(n_clusters))
    w = np.ones(n, dtype=dtype)
    center_id = rs.choice(n, p=w / w.sum())
    indices = np.full(n_clusters, -1, dtype=int)
    indices[0] = center_id
    closest = D[center_id].copy()
    current_pot = closest @ w
    for c in range(1, n_clusters):
        rand_vals = rs.uniform(size=n_local_trials) * current_pot
        cand = np.searchsorted(np.cumsum(w * closest), rand_vals)
        np.clip(cand, None, n - 1, out=cand)
        dist = np.minimum(closest[None, :], D[cand])
        pots = dist @ w.reshape(-1, 1)
        best = int(np.argmin(pots))
        current_pot = pots[best]
        closest = dist[best]
        indices[c] = cand[best]
    return indices


def greedy_density_balance(densities, labels, cand_off, num_class, axis, prior, bandwidth, n_select):
    """CRB stage 3 on the device (crb_sampling.py:264-338). densities/labels: stacked CUDA tensors of the candidates'
    per-box density and 1-based label, cand_off (n_cand+1). Returns the picked candidate positions (python list)."""
    from . import ops
    dev = densities.device
    order, scores = ops.kde_greedy(densities, labels, cand_off, num_class, axis.to(dev), normalise_prior(prior).to(dev),
                                   float(bandwidth), int(n_select))
    order = order.cpu().tolist()
    if any(o < 0 for o in order):
        raise RuntimeError("CRB stage 3: no candidate beat the initial best (-1) - the reference fails here as well")
    return order, scores.cpu()
