"""Sparse-conv oracle: rulebook + gather/mm/scatter conv + dense conv3d cross-check (TEST INFRASTRUCTURE ONLY).

Restates the behaviour of spconv-cu113==2.1.21 (third-party, NOT in /root/reference -> PARITY UNPINNED at the library
boundary) as exercised by pcdet/models/backbones_3d/spconv_backbone.py:69-180: SubMConv3d / SparseConv3d with weight
layout [C_out, kz, ky, kx, C_in] (pcdet/models/detectors/detector3d_template.py:455-484), kernel offset id
k = (kz*KY + ky)*KX + kx, pair (i -> o) at offset k iff p_in = p_out*stride - pad + k*dil.
Strided-conv output rows are in ascending linear (b,z,y,x) order (spconv GPU order).
`dense_conv_reference` pins all of it independently through torch.nn.functional.conv3d.
"""
import numpy as np
import torch
import torch.nn.functional as F


def _t3(v):
    return tuple(int(x) for x in (v if isinstance(v, (list, tuple, np.ndarray)) else (v, v, v)))


def out_shape(in_shape, ksize, stride, padding, dilation=(1, 1, 1)):
    return [int((in_shape[j] + 2 * padding[j] - dilation[j] * (ksize[j] - 1) - 1) // stride[j] + 1) for j in range(3)]


def _key(coords, shape):
    c = coords.astype(np.int64)
    return ((c[:, 0] * shape[0] + c[:, 1]) * shape[1] + c[:, 2]) * shape[2] + c[:, 3]


def _offsets(ksize):
    kz, ky, kx = ksize
    return [(z, y, x) for z in range(kz) for y in range(ky) for x in range(kx)]


def subm_rulebook(coords, spatial_shape, ksize, dilation=(1, 1, 1)):
    """nbr (K, N) int32: input row feeding output row o (== input row o's site) through offset k, else -1."""
    coords = np.asarray(coords, dtype=np.int32)
    ksize, dilation = _t3(ksize), _t3(dilation)
    n = len(coords)
    keys = _key(coords, spatial_shape)
    order = np.argsort(keys, kind="stable")
    skeys = keys[order]
    pad = [(ksize[j] // 2) * dilation[j] for j in range(3)]
    nbr = np.full((len(_offsets(ksize)), n), -1, dtype=np.int32)
    for k, off in enumerate(_offsets(ksize)):
        q = coords.astype(np.int64).copy()
        for j in range(3):
            q[:, 1 + j] = q[:, 1 + j] - pad[j] + off[j] * dilation[j]
        ok = np.ones(n, dtype=bool)
        for j in range(3):
            ok &= (q[:, 1 + j] >= 0) & (q[:, 1 + j] < spatial_shape[j])
        qk = _key(np.where(ok[:, None], q, 0), spatial_shape)
        pos = np.searchsorted(skeys, qk)
        pos_c = np.minimum(pos, max(n - 1, 0))
        hit = ok & (n > 0) & (skeys[pos_c] == qk)
        nbr[k, hit] = order[pos_c[hit]]
    return nbr


def sparse_rulebook(coords, batch_size, in_shape, ksize, stride, padding, dilation=(1, 1, 1)):
    """(out_coords (M,4) ascending key, out_shape, nbr (K,M), nbr_t (K,N))."""
    coords = np.asarray(coords, dtype=np.int32)
    ksize, stride, padding, dilation = _t3(ksize), _t3(stride), _t3(padding), _t3(dilation)
    oshape = out_shape(in_shape, ksize, stride, padding, dilation)
    n = len(coords)
    K = len(_offsets(ksize))
    cand_key = np.full((K, n), -1, dtype=np.int64)
    for k, off in enumerate(_offsets(ksize)):
        q = coords.astype(np.int64).copy()
        ok = np.ones(n, dtype=bool)
        for j in range(3):
            v = q[:, 1 + j] + padding[j] - off[j] * dilation[j]
            ok &= (v >= 0) & (v % stride[j] == 0)
            v = v // stride[j]
            ok &= v < oshape[j]
            q[:, 1 + j] = v
        kk = _key(np.where(ok[:, None], q, 0), oshape)
        cand_key[k, ok] = kk[ok]
    uniq = np.unique(cand_key[cand_key >= 0])
    m = len(uniq)
    out_coords = np.zeros((m, 4), dtype=np.int32)
    r = uniq.copy()
    out_coords[:, 3] = r % oshape[2]; r //= oshape[2]
    out_coords[:, 2] = r % oshape[1]; r //= oshape[1]
    out_coords[:, 1] = r % oshape[0]; r //= oshape[0]
    out_coords[:, 0] = r
    nbr = np.full((K, m), -1, dtype=np.int32)
    nbr_t = np.full((K, n), -1, dtype=np.int32)
    for k in range(K):
        ok = cand_key[k] >= 0
        o = np.searchsorted(uniq, cand_key[k, ok])
        i = np.nonzero(ok)[0]
        nbr[k, o] = i
        nbr_t[k, i] = o
    return out_coords, oshape, nbr, nbr_t


def pairs_from_table(nbr):
    """spconv-format indice_pairs [2,K,Nmax] (-1 padded) + indice_pair_num [K]; pairs ascending in output row."""
    K, m = nbr.shape
    pairs = np.full((2, K, m), -1, dtype=np.int32)
    num = np.zeros((K,), dtype=np.int32)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        num[k] = len(o)
        pairs[0, k, :len(o)] = nbr[k, o]
        pairs[1, k, :len(o)] = o
    return pairs, num


def conv_forward(feat, nbr, weight, dtype=torch.float32):
    """spconv 'Native' algorithm: out = sum_k index_add(out_rows_k, feat[in_rows_k] @ W_k^T). weight [Cout,kz,ky,kx,Cin]."""
    feat = torch.as_tensor(feat).to(dtype)
    w = torch.as_tensor(weight).to(dtype)
    cout, cin = w.shape[0], w.shape[-1]
    w = w.reshape(cout, -1, cin)
    K, m = nbr.shape
    out = torch.zeros((m, cout), dtype=dtype)
    for k in range(K):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o) == 0:
            continue
        i = torch.as_tensor(nbr[k, o].astype(np.int64))
        out.index_add_(0, torch.as_tensor(o.astype(np.int64)), feat.index_select(0, i) @ w[:, k, :].t())
    return out


def conv_backward(feat, nbr, weight, dout, dtype=torch.float32):
    """(dX, dW) of conv_forward by the pair-wise definition (SURVEY.md 2.4 'Backward')."""
    feat = torch.as_tensor(feat).to(dtype)
    dout = torch.as_tensor(dout).to(dtype)
    w = torch.as_tensor(weight).to(dtype)
    shape = w.shape
    cout, cin = w.shape[0], w.shape[-1]
    w = w.reshape(cout, -1, cin)
    dx = torch.zeros_like(feat)
    dw = torch.zeros_like(w)
    for k in range(nbr.shape[0]):
        o = np.nonzero(nbr[k] >= 0)[0]
        if len(o) == 0:
            continue
        i = torch.as_tensor(nbr[k, o].astype(np.int64))
        ot = torch.as_tensor(o.astype(np.int64))
        g = dout.index_select(0, ot)
        dx.index_add_(0, i, g @ w[:, k, :])
        dw[:, k, :] = g.t() @ feat.index_select(0, i)
    return dx, dw.reshape(shape)


def dense(feat, coords, batch_size, spatial_shape):
    """SparseConvTensor.dense(): (B, C, D, H, W) (height_compression.py:21)."""
    feat = torch.as_tensor(feat)
    c = torch.as_tensor(np.asarray(coords)).long()
    out = torch.zeros((batch_size, feat.shape[1], *[int(s) for s in spatial_shape]), dtype=feat.dtype)
    out[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]] = feat
    return out


def dense_conv_reference(feat, coords, batch_size, in_shape, weight, stride, padding, dilation=(1, 1, 1), subm=False,
                         dtype=torch.float64):
    """Independent check: dense conv3d over the scattered input, read back at the active output sites.
    Returns (out_coords ascending key (or the input coords for subm), out_feats)."""
    w = torch.as_tensor(weight).to(dtype)
    x = dense(torch.as_tensor(feat).to(dtype), coords, batch_size, in_shape)
    y = F.conv3d(x, w.permute(0, 4, 1, 2, 3).contiguous(), stride=_t3(stride), padding=_t3(padding), dilation=_t3(dilation))
    if subm:
        oc = np.asarray(coords)
    else:
        occ = dense(torch.ones((len(coords), 1), dtype=dtype), coords, batch_size, in_shape)
        hit = F.conv3d(occ, torch.ones((1, 1, *w.shape[1:4]), dtype=dtype), stride=_t3(stride), padding=_t3(padding),
                       dilation=_t3(dilation))
        oc = torch.nonzero(hit[:, 0] > 0.5).numpy().astype(np.int32)  # nonzero() is lexicographic == ascending key
    c = torch.as_tensor(oc).long()
    return oc, y[c[:, 0], :, c[:, 1], c[:, 2], c[:, 3]]
