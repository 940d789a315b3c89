"""Selection primitives of the other acquisition strategies of the reference, on the device without per-pick host syncs
(SURVEY.md 8f item 3). They consume the same detector outputs as CRB; the strategy classes around them (data loading,
label bookkeeping) are the reference's control plane and stay where they are.

  furthest_first    pcdet/query_strategies/coreset_sampling.py:12-52  (CORESET: greedy k-centre on embeddings)
  kmeans_pp_select  pcdet/query_strategies/badge_sampling.py:190-196  (BADGE: k-means++ seeding on gradient embeddings;
                    same sklearn RandomState(0) stream as CRB stage 2, see crb_host.kmeans_plusplus_indices)
"""
import torch

from . import crb_host, ops


def pairwise_squared_distances(x, y):
    """coreset_sampling.py:12-29: ||x||^2 + ||y||^2 - 2 x y^T in fp32, NaN -> 0, clamped at 0."""
    n, m = x.shape[0], y.shape[0]
    x, y = x.reshape(n, -1).float(), y.reshape(m, -1).float()
    dist = (x ** 2).sum(1).view(n, 1) + (y ** 2).sum(1).view(1, m) - 2.0 * torch.mm(x, y.t().contiguous())
    dist = torch.where(dist != dist, torch.zeros((), device=dist.device), dist)
    return torch.clamp(dist, 0.0, float("inf"))


def furthest_first(X, X_set, n):
    """coreset_sampling.py:31-52. Starts from the MEAN distance to the labelled set (the reference's `min_dist =
    dist_ctr.mean(1)`), then n greedy picks: arg-max, then element-wise min with the distance to the new centre. The
    reference updates `min_dist` with a Python loop over all m elements per pick; on CUDA tensors the whole loop runs in
    crb3d_furthest_first (distance to the new centre + running minimum + next arg-max in one pass per pick); the index never
    visits the host. CUDA tensors only (no CPU path). Returns a LongTensor (n,) on X's device."""
    m = X.shape[0]
    X = X.reshape(m, -1).float()
    min_dist = pairwise_squared_distances(X, X_set).mean(1)
    # the greedy loop as two kernel launches per pick (csrc/data_prims.cu), no per-pick host traffic; CUDA tensors only
    return ops.furthest_first(X.contiguous(), min_dist.contiguous(), n)


def kmeans_pp_select(embeddings, n, seed=0):
    """badge_sampling.py:190-196: indices of sklearn.cluster.kmeans_plusplus(embeddings, n, random_state=seed) - computed by
    the restated seeding of crb_host (device distances, sklearn's RandomState stream)."""
    D = ops.pairwise_sqdist(embeddings.reshape(embeddings.shape[0], -1))          # fp64 Gram tiles on the device
    return crb_host.kmeans_plusplus_indices(D.cpu().numpy(), n, seed=seed)
