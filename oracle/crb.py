"""CRB acquisition oracle (TEST INFRASTRUCTURE ONLY).

Runs the reference's own library calls, restating pcdet/query_strategies/crb_sampling.py:
  stage 1  :86-100,119-121  torch.unique + Categorical(probs).entropy(), stable ascending sort then reverse
  stage 2  :219-226         sklearn.cluster.kmeans_plusplus(X, n_clusters, random_state=0)
  stage 3  :247-338         uniform prior per class on np.linspace(-50, int(max)+50, 400); greedy loop with
                            sklearn KernelDensity(gaussian, bandwidth).score_samples + scipy.stats.entropy
"""
import numpy as np
import scipy.stats
import torch
from scipy.stats import uniform
from sklearn.cluster import kmeans_plusplus
from sklearn.neighbors import KernelDensity
from torch.distributions import Categorical


def label_entropy(pred_labels, num_class):
    """pred_labels: 1-based int labels of one frame. Returns python 0 for an empty frame (crb_sampling.py:87-88)."""
    labels = torch.as_tensor(np.asarray(pred_labels)).long()
    value, counts = torch.unique(labels, return_counts=True)
    if len(value) == 0:
        return 0.0
    unique_proportions = torch.ones(num_class)
    unique_proportions[value - 1] = counts.float()
    return float(Categorical(probs=unique_proportions / sum(counts)).entropy())


def stage1_shortlist(frame_ids, entropies, k):
    """dict(sorted(items, key=value)) ascending (stable), reversed, first k (crb_sampling.py:119-121)."""
    items = sorted(zip(frame_ids, entropies), key=lambda it: it[1])
    return [f for f, _ in items][::-1][:k]


def kmeanspp_indices(X, n_clusters):
    _, idx = kmeans_plusplus(np.asarray(X), n_clusters=n_clusters, random_state=0)
    return idx


def build_prior(density_all, label_all, num_class, alpha=0.95):
    """crb_sampling.py:252-260. density_all float32 (n,), label_all int (n,) 1-based. Returns (x_axis, uniform pdf) lists."""
    density_all = torch.as_tensor(np.asarray(density_all, dtype=np.float32))
    label_all = torch.as_tensor(np.asarray(label_all)).long()
    unique_labels, label_counts = torch.unique(label_all, return_counts=True)
    sorted_density = [torch.sort(density_all[label_all == u])[0] for u in unique_labels]
    gmax = [int(sorted_density[u][-1]) for u in range(len(unique_labels))]
    ghigh = [int(sorted_density[u][int(alpha * label_counts[u])]) for u in range(len(unique_labels))]
    glow = [int(sorted_density[u][-int(alpha * label_counts[u])]) for u in range(len(unique_labels))]
    x_axis = [np.linspace(-50, int(gmax[i]) + 50, 400) for i in range(num_class)]
    prior = [uniform.pdf(x_axis[i], glow[i], ghigh[i] - glow[i]) for i in range(num_class)]
    return x_axis, prior


def greedy_density_balance(density_list, label_list, x_axis, prior, num_class, select_nums, bandwidth=5):
    """crb_sampling.py:264-338 with python lists of per-frame float32 density arrays / int label arrays.
    Returns (picked candidate positions in the ORIGINAL list, best inverse_coff per round (nan for round 0))."""
    density_list = [np.asarray(d, dtype=np.float32) for d in density_list]
    label_list = [np.asarray(l) for l in label_list]
    ids = list(range(len(density_list)))
    sel_d = np.zeros((0,), np.float32)
    sel_l = np.zeros((0,), np.int64)
    picked, scores = [], []
    for j in range(select_nums):
        if j == 0:
            best_i, best = 0, float("nan")
        else:
            best_i, best = None, -1
            for i in range(len(density_list)):
                props = np.zeros(num_class)
                for cls in range(num_class):
                    if (label_list[i] == cls + 1).sum() == 0:
                        props[cls] = 1
                    else:
                        d = np.concatenate([sel_d[sel_l == cls + 1], density_list[i][label_list[i] == cls + 1]])
                        kde = KernelDensity(kernel="gaussian", bandwidth=bandwidth).fit(d[:, None])
                        logprob = kde.score_samples(x_axis[cls][:, None])
                        kl = scipy.stats.entropy(prior[cls], np.exp(logprob))
                        props[cls] = 2 / np.pi * np.arctan(np.pi / 2 * kl)
                inv = np.mean(1 - props)
                if inv > best:
                    best, best_i = inv, i
            if best_i is None:
                raise RuntimeError("no candidate beat the initial best (-1)")
        sel_d = np.concatenate([sel_d, density_list[best_i]])
        sel_l = np.concatenate([sel_l, label_list[best_i]])
        picked.append(ids[best_i])
        scores.append(best)
        del density_list[best_i], label_list[best_i], ids[best_i]
    return picked, scores
