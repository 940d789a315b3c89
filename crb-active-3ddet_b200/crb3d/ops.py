"""Torch-tensor front end of the C ABI (include/crb3d.h). PyTorch is used for device memory and streams only;
every function here launches hand-written sm_100a kernels from libcrb3d_sm100.so on the current CUDA stream.
There is no CPU path: CPU tensors are rejected (except for the two explicit *_cpu host ops).
"""
import ctypes
import os
from ctypes import c_size_t, byref

import numpy as np
import torch

from . import _lib

_I3 = ctypes.c_int * 3
_F3 = ctypes.c_float * 3
_F6 = ctypes.c_float * 6


def _i3(v):
    v = [int(x) for x in (v if isinstance(v, (list, tuple, np.ndarray, torch.Size)) else [v, v, v])]
    assert len(v) == 3
    return _I3(*v)


def _p(t):
    if t is None:
        return None
    return ctypes.c_void_p(t.data_ptr())


def _stream(dev=None):
    st = torch.cuda.current_stream(dev)
    if st.device_index not in _lib._DIAG_DEVICES:   # first launch on this device: hook up the bounded-wait diagnostics
        _lib.init_device(st.device_index)
    return ctypes.c_void_p(st.cuda_stream)


def _need_cuda(*ts):
    for t in ts:
        if t is not None and not t.is_cuda:
            raise RuntimeError("crb3d ops run on CUDA tensors only (no CPU fallback)")


def _check(**named):
    """CHECK_INPUT of the reference's pybind wrappers (e.g. pointnet2_stack/src/ball_query.cpp:14-17: CUDA + contiguous) plus
    the dtype the kernel reads: the C ABI takes raw addresses, so a strided view, an int64 count vector or a half tensor would
    silently produce garbage. name=(tensor, dtype); None tensors are skipped."""
    for name, (t, dtype) in named.items():
        if t is None:
            continue
        if not t.is_cuda:
            raise RuntimeError("crb3d: %s must be a CUDA tensor (no CPU fallback)" % name)
        if t.dtype != dtype:
            raise TypeError("crb3d: %s must be %s, got %s" % (name, dtype, t.dtype))
        if not t.is_contiguous():
            raise ValueError("crb3d: %s must be contiguous" % name)


def _check_len(**named):
    """name=(tensor, expected number of elements): count / offset vectors are indexed by the batch size on the device."""
    for name, (t, n) in named.items():
        if t is not None and t.numel() != n:
            raise ValueError("crb3d: %s must hold %d entries, got %d" % (name, n, t.numel()))


def _f32c(t):
    return t if (t.dtype == torch.float32 and t.is_contiguous()) else t.float().contiguous()


def _i32c(t):
    return t if (t.dtype == torch.int32 and t.is_contiguous()) else t.int().contiguous()


def debug_mark(slot, value):
    """Debug: queues a one-thread kernel that stores `value` in host-visible marker `slot` (see csrc/diag.cu)."""
    lib = _lib.load()
    lib.crb3d_debug_mark.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
    _lib.check(lib.crb3d_debug_mark(int(slot), int(value), _stream()), "crb3d_debug_mark")


def debug_read_markers(n=16):
    lib = _lib.load()
    out = (ctypes.c_int * n)()
    lib.crb3d_debug_read_markers.argtypes = [ctypes.c_void_p, ctypes.c_int]
    _lib.check(lib.crb3d_debug_read_markers(out, n), "crb3d_debug_read_markers")
    return list(out)


_WS = {}
WS_TAG = None   # extra key: every captured copy of the whole-step graph owns its scratch (copies replay concurrently)


def _ws(nbytes, device):
    """Grow-only scratch buffer per (device, stream, WS_TAG) from torch's caching allocator."""
    key = (device.index if device.index is not None else torch.cuda.current_device(),
           torch.cuda.current_stream(device).cuda_stream, WS_TAG)
    buf = _WS.get(key)
    if buf is None or buf.numel() < nbytes:
        buf = torch.empty(max(int(nbytes), 1 << 20), dtype=torch.uint8, device=device)
        _WS[key] = buf
    return buf


def drop_workspaces(tags):
    """Frees the scratch buffers of graph copies that are being discarded (keys carry the copy's WS_TAG)."""
    tags = set(tags)
    for key in [k for k in _WS if k[2] in tags]:
        del _WS[key]


def _ws_bytes(fn, *args):
    n = c_size_t(0)
    _lib.call(fn, *args, byref(n))
    return int(n.value)


# ----------------------------------------------------------------------------------------------- voxelize
def voxelize(points, frame_offsets, batch_size, pc_range, voxel_size, max_pts, max_voxels, xyz_col=0, feat_col=0,
             n_feat=None, want_voxels=False, want_mean=True, sync=True):
    """Hard voxelization + MeanVFE. points (N, S) f32 CUDA; frame_offsets (B+1) int32 CUDA.
    Returns dict(mean (M,C), voxels (M,P,C) | None, coords (M,4) i32 [b,z,y,x], num_points (M), frame_voxel_offsets (B+1)).
    One host sync (reads M) unless sync=False (capacity-sized outputs + device-side count, CUDA-graph capturable; rows of
    `points` at or beyond frame_offsets[B] are ignored)."""
    _need_cuda(points, frame_offsets)
    points = _f32c(points)
    frame_offsets = _i32c(frame_offsets)
    if frame_offsets.numel() != batch_size + 1:      # the kernels index frame_offsets[0 .. batch_size]
        raise ValueError("crb3d: frame_offsets must hold batch_size + 1 = %d entries, got %d" % (batch_size + 1, frame_offsets.numel()))
    n, stride = points.shape
    n_feat = stride - feat_col if n_feat is None else n_feat
    pc_range = [float(x) for x in pc_range]
    voxel_size = [float(x) for x in voxel_size]
    grid = [int(round((pc_range[3 + j] - pc_range[j]) / voxel_size[j])) for j in range(3)]
    cap = max(1, min(n, batch_size * max_voxels))
    dev = points.device
    mean = torch.empty((cap, n_feat), dtype=torch.float32, device=dev) if want_mean else None
    voxels = torch.empty((cap, max_pts, n_feat), dtype=torch.float32, device=dev) if want_voxels else None
    coords = torch.empty((cap, 4), dtype=torch.int32, device=dev)
    num = torch.empty((cap,), dtype=torch.int32, device=dev)
    voff = torch.empty((batch_size + 1,), dtype=torch.int32, device=dev)
    wsb = _ws_bytes("crb3d_voxelize_workspace_bytes", n, batch_size, max_pts)
    ws = _ws(wsb, dev)
    _lib.call("crb3d_voxelize", _p(points), n, stride, xyz_col, feat_col, n_feat, _p(frame_offsets), batch_size,
              _F6(*pc_range), _F3(*voxel_size), _I3(*grid), max_pts, max_voxels, _p(mean), _p(voxels), _p(coords),
              _p(num), _p(voff), _p(ws), ws.numel(), _stream(dev))
    if not sync:   # static mode: capacity-sized outputs, the voxel count stays on the device (frame_voxel_offsets[B])
        return dict(mean=mean, voxels=voxels, coords=coords, num_points=num, frame_voxel_offsets=voff, grid_size=grid,
                    n_dev=voff[batch_size:])
    m = int(voff[-1].item())
    return dict(mean=mean[:m] if mean is not None else None, voxels=voxels[:m] if voxels is not None else None,
                coords=coords[:m], num_points=num[:m], frame_voxel_offsets=voff, grid_size=grid)


# ----------------------------------------------------------------------------------------------- rulebook
def conv_out_shape(in_shape, ksize, stride, padding, dilation=(1, 1, 1)):
    out = _I3(0, 0, 0)
    _lib.call("crb3d_conv_out_shape", _i3(in_shape), _i3(ksize), _i3(stride), _i3(padding), _i3(dilation), out)
    return [out[0], out[1], out[2]]


class CellMap(object):
    """cell -> row map of a level produced by a strided rulebook: the output-cell bitmap and its word ranks, two views into
    that call's private workspace (kept alive here). Valid for `coords` (the rows in rank order) and `shape` only."""

    def __init__(self, ws, batch_size, shape, coords):
        b_off, r_off, nw = c_size_t(0), c_size_t(0), ctypes.c_int64(0)
        _lib.call("crb3d_sparse_rulebook_cellmap", int(batch_size), _i3(shape), byref(b_off), byref(r_off), byref(nw))
        self.ws, self.shape, self.coords = ws, [int(v) for v in shape], coords
        self.bitmap = ws.data_ptr() + int(b_off.value)
        self.rank = ws.data_ptr() + int(r_off.value)

    def matches(self, coords, shape):
        return coords.data_ptr() == self.coords.data_ptr() and coords.shape[0] == self.coords.shape[0] and \
            [int(v) for v in shape] == self.shape


def subm_rulebook(coords, spatial_shape, ksize, dilation=(1, 1, 1), n_dev=None, cellmap=None):
    """Neighbour table (K, n) int32 for a submanifold conv (output rows == input rows). n_dev: device int32 row count when
    `coords` is a capacity-sized static buffer (rows beyond it are not written). cellmap: the CellMap of the strided rulebook
    that produced `coords` - the table is then read off its bitmap ranks instead of building and probing a hash table."""
    _need_cuda(coords)
    coords = _i32c(coords)
    n = coords.shape[0]
    k = _i3(ksize)
    K = k[0] * k[1] * k[2]
    nbr = torch.empty((K, n), dtype=torch.int32, device=coords.device)
    if cellmap is not None and cellmap.matches(coords, spatial_shape):
        _lib.call("crb3d_subm_rulebook_cellmap", _p(coords), n, _p(n_dev), _i3(spatial_shape), k, _i3(dilation),
                  ctypes.c_void_p(cellmap.bitmap), ctypes.c_void_p(cellmap.rank), _p(nbr), _stream(coords.device))
        return nbr
    ws = _ws(_ws_bytes("crb3d_subm_rulebook_workspace_bytes", n), coords.device)
    _lib.call("crb3d_subm_rulebook", _p(coords), n, _p(n_dev), _i3(spatial_shape), k, _i3(dilation), _p(nbr), _p(ws), ws.numel(),
              _stream(coords.device))
    return nbr


def sparse_rulebook(coords, batch_size, in_shape, ksize, stride, padding, dilation=(1, 1, 1), want_transpose=True, want_cellmap=False):
    """Strided sparse conv rulebook. Returns (out_coords (n_out,4) ascending key order, out_shape, nbr (K,n_out),
    nbr_t (K,n_in) | None[, CellMap of the output level]). One host sync (reads n_out)."""
    _need_cuda(coords)
    coords = _i32c(coords)
    dev = coords.device
    n_in = coords.shape[0]
    out_shape = conv_out_shape(in_shape, ksize, stride, padding, dilation)
    k = _i3(ksize)
    K = k[0] * k[1] * k[2]
    wsb = _ws_bytes("crb3d_sparse_rulebook_workspace_bytes", batch_size, _i3(out_shape))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)  # private: must survive between the two phases
    cells = batch_size * out_shape[0] * out_shape[1] * out_shape[2]
    cap = max(1, min(cells, n_in * K))
    # most strided layers produce about as many outputs as inputs; retry with the true count if the guess is small
    guess = max(1, min(cap, 2 * n_in + 1024))
    n_out_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    out_coords = torch.empty((guess, 4), dtype=torch.int32, device=dev)
    args = (_p(coords), n_in, None, batch_size, _i3(in_shape), _i3(out_shape), k, _i3(stride), _i3(padding), _i3(dilation))
    _lib.call("crb3d_sparse_rulebook_coords", *args, _p(out_coords), guess, _p(n_out_dev), _p(ws), wsb, _stream(dev))
    n_out = int(n_out_dev.item())
    if n_out > guess:
        out_coords = torch.empty((n_out, 4), dtype=torch.int32, device=dev)
        _lib.call("crb3d_sparse_rulebook_coords", *args, _p(out_coords), n_out, _p(n_out_dev), _p(ws), wsb, _stream(dev))
    out_coords = out_coords[:n_out]
    nbr = torch.empty((K, n_out), dtype=torch.int32, device=dev)
    nbr_t = torch.empty((K, n_in), dtype=torch.int32, device=dev) if want_transpose else None
    _lib.call("crb3d_sparse_rulebook_pairs", *args, n_out, _p(nbr), _p(nbr_t), _p(ws), wsb, _stream(dev))
    if want_cellmap:
        return out_coords, out_shape, nbr, nbr_t, CellMap(ws, batch_size, out_shape, out_coords)
    return out_coords, out_shape, nbr, nbr_t


def sparse_rulebook_static(coords, n_in_dev, batch_size, in_shape, ksize, stride, padding, cap_out, dilation=(1, 1, 1),
                           want_cellmap=False):
    """Strided sparse conv rulebook without any host synchronisation (CUDA-graph capturable): `coords` (cap_in, 4) holds
    n_in_dev[0] valid rows; returns (out_coords (cap_out, 4), out_shape, nbr (K, cap_out), n_out_dev (1,) int32 device).
    n_out_dev is the TRUE count - if it exceeds cap_out the extra outputs were dropped (callers check it afterwards)."""
    _need_cuda(coords, n_in_dev)
    coords = _i32c(coords)
    dev = coords.device
    cap_in = coords.shape[0]
    out_shape = conv_out_shape(in_shape, ksize, stride, padding, dilation)
    k = _i3(ksize)
    K = k[0] * k[1] * k[2]
    wsb = _ws_bytes("crb3d_sparse_rulebook_workspace_bytes", batch_size, _i3(out_shape))
    ws = torch.empty(wsb, dtype=torch.uint8, device=dev)
    n_out_dev = torch.zeros(1, dtype=torch.int32, device=dev)
    out_coords = torch.zeros((cap_out, 4), dtype=torch.int32, device=dev)
    nbr = torch.empty((K, cap_out), dtype=torch.int32, device=dev)
    args = (_p(coords), cap_in, _p(n_in_dev), batch_size, _i3(in_shape), _i3(out_shape), k, _i3(stride), _i3(padding), _i3(dilation))
    _lib.call("crb3d_sparse_rulebook_coords", *args, _p(out_coords), cap_out, _p(n_out_dev), _p(ws), wsb, _stream(dev))
    _lib.call("crb3d_sparse_rulebook_pairs", *args, cap_out, _p(nbr), None, _p(ws), wsb, _stream(dev))
    if want_cellmap:
        return out_coords, out_shape, nbr, n_out_dev, CellMap(ws, batch_size, out_shape, out_coords)
    return out_coords, out_shape, nbr, n_out_dev


def compact_pairs(nbr):
    """spconv-format (indice_pairs [2,K,n_out] padded with -1, indice_pair_num [K]) from a neighbour table."""
    _need_cuda(nbr)
    K, n_out = nbr.shape
    dev = nbr.device
    pairs = torch.full((2, K, max(n_out, 1)), -1, dtype=torch.int32, device=dev)
    num = torch.zeros((K,), dtype=torch.int32, device=dev)
    ws = _ws(_ws_bytes("crb3d_rulebook_compact_pairs_workspace_bytes", K, n_out), dev)
    _lib.call("crb3d_rulebook_compact_pairs", _p(nbr), K, n_out, max(n_out, 1), _p(pairs[0]), _p(pairs[1]), _p(num),
              _p(ws), ws.numel(), _stream(dev))
    return pairs[:, :, :n_out] if n_out > 0 else pairs[:, :, :0], num


# ----------------------------------------------------------------------------------------------- sparse conv
def spconv_forward(feat, nbr, weight, scale=None, shift=None, relu=False, transpose=False, kmap=None, tf32=None, n_dev=None,
                   round_out=False):
    """out[o] = sum_k feat[nbr[k][o]] @ W_k. weight: [C_out, (kz,ky,kx)|K, C_in] (spconv layout).
    transpose=True computes the input gradient: feat is dY (rows of the conv OUTPUT), nbr the transposed table."""
    _need_cuda(feat, nbr, weight)
    feat = _f32c(feat)
    nbr = _i32c(nbr)
    weight = _f32c(weight)
    K, n_out = nbr.shape
    cout_w, cin_w = weight.shape[0], weight.shape[-1]
    assert weight.numel() == cout_w * K * cin_w, "weight does not match the rulebook's kernel volume"
    if not transpose:
        cin, cout = cin_w, cout_w
        strides = (K * cin_w, cin_w, 1)
    else:
        cin, cout = cout_w, cin_w
        strides = (1, cin_w, K * cin_w)
    assert feat.shape[1] == cin, (feat.shape, cin)
    # n_dev: device row count of a capacity-sized table; rows beyond it are left untouched (no consumer reads them)
    out = torch.empty((n_out, cout), dtype=torch.float32, device=feat.device)
    prof = PROFILE
    if prof is not None and prof["mode"] == "time":
        e0 = torch.cuda.Event(enable_timing=True)
        e1 = torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(feat.device))
    use_tc = ((SPCONV_TF32 if tf32 is None else tf32) and K <= 27 and cin in (4, 8, 16, 32, 64) and cout in (16, 32, 64, 128)
              and feat.shape[0] > 0)
    if use_tc:
        # tcgen05 path wants K-major weight rows [C_out', K, C_in']: the forward layout as is, its transpose for dX
        w_tc = weight.reshape(cout_w, K, cin_w)
        if transpose:
            w_tc = w_tc.permute(2, 1, 0).contiguous()
        _lib.call("crb3d_spconv_forward_tf32", _p(feat), feat.shape[0], _p(nbr), _p(w_tc), n_out, K, cin, cout, _p(kmap),
                  _p(_f32c(scale)) if scale is not None else None, _p(_f32c(shift)) if shift is not None else None,
                  int(bool(relu)) | (2 if round_out else 0) | (0 if SPCONV_GROUPED else 4), _p(out), _p(n_dev), _stream(feat.device))
    else:
        _lib.call("crb3d_spconv_forward_f32", _p(feat), _p(nbr), _p(weight), n_out, K, cin, cout, strides[0], strides[1],
                  strides[2], _p(kmap), _p(_f32c(scale)) if scale is not None else None,
                  _p(_f32c(shift)) if shift is not None else None, int(bool(relu)), _p(out), _p(n_dev), _stream(feat.device))
    if prof is not None:
        if prof["mode"] == "time":
            e1.record(torch.cuda.current_stream(feat.device))
            prof["records"].append((e0, e1))
        else:  # "pairs": algorithmic work of this launch (host sync; only used outside timed regions)
            prof["records"].append(dict(n_out=n_out, K=K, cin=cin, cout=cout, pairs=int((nbr >= 0).sum().item())))
    return out


# tensor-core (tcgen05, TF32 inputs / fp32 accumulate) sparse conv for the layers it covers; False = exact-fp32 SIMT path
SPCONV_TF32 = False
SPCONV_GROUPED = True     # C_in <= 8 (the input layer) on the grouped-stage kernel (csrc/spconv_tc_grp.cu); False: one offset per stage

# bench.py hook: None, or {"mode": "time" | "pairs", "records": []} (see bench.py roofline section)
PROFILE = None


def spconv_wgrad(feat, dout, nbr, weight_shape, accumulate_into=None):
    """dW in the spconv layout for out = conv(feat, nbr, W): dW[co,k,ci] = sum_o dY[o,co] * feat[nbr[k][o],ci]."""
    _need_cuda(feat, dout, nbr)
    feat = _f32c(feat)
    dout = _f32c(dout)
    nbr = _i32c(nbr)
    K, n_out = nbr.shape
    cin, cout = feat.shape[1], dout.shape[1]
    dw = accumulate_into if accumulate_into is not None else torch.empty(weight_shape, dtype=torch.float32, device=feat.device)
    assert dw.is_contiguous() and dw.numel() == cout * K * cin
    ws = _ws(_ws_bytes("crb3d_spconv_wgrad_workspace_bytes", n_out, K, cin, cout), feat.device)
    _lib.call("crb3d_spconv_wgrad_f32", _p(feat), _p(dout), _p(nbr), n_out, K, cin, cout,
              int(accumulate_into is not None), _p(dw), _p(ws), ws.numel(), _stream(feat.device))
    return dw


# ----------------------------------------------------------------------------------------------- BEV GEMMs (tcgen05)
def round_tf32(t):
    """fp32 -> nearest TF32 value (10 mantissa bits), kept in fp32 storage. The tensor cores truncate fp32 operands;
    weights rounded once here (and activations rounded by the producing epilogue) are read exactly instead."""
    i = t.detach().contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def tf32_weight(conv):
    """A conv module's weight rounded to TF32 (round to nearest) once per parameter version: the tensor core truncates fp32
    operands, a systematic shrink of ~3e-4 per operand per layer (the BEV stack does the same in its inference plan). Used by
    both inference routes of the sparse convs (module forward and the captured static step) so that they stay bit-identical."""
    w = conv.weight
    key = (w._version, w.data_ptr())
    c = getattr(conv, "_crb3d_w_tf32", None)
    if c is None or c[0] != key:
        c = (key, round_tf32(w))
        conv._crb3d_w_tf32 = c
    return c[1]


# opt-in: thread-block clusters with TMA multicast of shared operand k-blocks in the long-K GEMMs (csrc/bev_gemm_tc.cu). Same
# results, measured SLOWER on B200 at the BEV shapes (tools/bench_gemm_cluster.py) - kept for A/B measurements
GEMM_CLUSTERS = False
GEMM_PAIR_STREAM = os.environ.get("CRB3D_GEMM_PAIR_STREAM", "1") == "1"   # pair GEMM: weights streamed with the activations (6 stages) instead of resident (4)
GEMM_PAIR_SHORTK = os.environ.get("CRB3D_GEMM_PAIR_SHORTK", "0") == "1"   # A/B: the K <= 128 deblock on the pair kernel too
GEMM_PAIRS = os.environ.get("CRB3D_GEMM_PAIRS", "1") != "0"      # long-K 256-column deblock on CTA pairs (csrc/bev_gemm_pair.cu); False = the single-CTA kernel (A/B runs)


def bev_gemm(a, weight, bias, relu, segs, n_sub=1, up=0, in_hw=(0, 0), round_out=False):
    """D = A @ W^T (+bias) (ReLU) on the tensor cores with fused output placement (csrc/bev_gemm_tc.cu).
    a: (M, K) fp32 CUDA, unit column stride (row stride = a.stride(0)); weight: contiguous (n_sub*N, K);
    bias: (N,) or None; segs: [(out_tensor, col_begin, width, row_stride_floats)] - column segment -> out_tensor's
    storage starting at its data_ptr(). up=2: the four (dy,dx) slices of a k=s=2 transposed conv, in_hw = input (H, W)."""
    _need_cuda(a, weight)
    assert a.dtype == torch.float32 and a.dim() == 2 and a.stride(1) == 1
    weight = _f32c(weight)
    M, K = a.shape
    N = weight.shape[0] // n_sub
    assert weight.shape[1] == K and 1 <= len(segs) <= 3
    ptrs = (ctypes.c_void_p * 3)(*[s[0].data_ptr() for s in segs], *([None] * (3 - len(segs))))
    cb = (ctypes.c_int * 3)(*[int(s[1]) for s in segs], *([0] * (3 - len(segs))))
    wd = (ctypes.c_int * 3)(*[int(s[2]) for s in segs], *([0] * (3 - len(segs))))
    rs = (ctypes.c_longlong * 3)(*[int(s[3]) for s in segs], *([0] * (3 - len(segs))))
    _lib.call("crb3d_bev_gemm_tf32", _p(a), M, K, a.stride(0), _p(weight), N, n_sub, _p(_f32c(bias)) if bias is not None else None,
              int(bool(relu)) | (2 if round_out else 0) | (4 if GEMM_CLUSTERS else 0) | (0 if GEMM_PAIRS else 8) | (0 if GEMM_PAIR_STREAM else 16) | (32 if GEMM_PAIR_SHORTK else 0), len(segs), ptrs, cb, wd, rs, int(up), int(in_hw[0]), int(in_hw[1]),
              _stream(a.device))


CONV_VARIANT = int(os.environ.get("CRB3D_CONV_VARIANT", "0"))   # bit 0: single-CTA kernel, bits 1-2: its experiments (bev_conv_tc.cu)


def pack_conv3x3_weight(weight, split=None):
    """(C_out, C_in, 3, 3) conv weight -> the slab layout of csrc/bev_conv_tc.cu:
    [C_out/128][tap = ky*3+kx][C_in/16][half][4 slabs][64 co][4 ci] (contiguous fp32; each CTA of a pair holds 64 output
    channels of a slice); split=False: [C_out/128][tap][C_in/16][4 slabs][128 co][4 ci] of the single-CTA kernel."""
    cout, cin = weight.shape[0], weight.shape[1]
    assert tuple(weight.shape[2:]) == (3, 3) and cout % 128 == 0 and cin % 16 == 0
    split = not (CONV_VARIANT & 1) if split is None else split
    w = round_tf32(weight.detach().float()).permute(0, 2, 3, 1)             # (cout, ky, kx, cin)
    if split:
        w = w.reshape(cout // 128, 2, 64, 9, cin // 16, 4, 4)                # nh, half, co, tap, chunk, slab, ci
        return w.permute(0, 3, 4, 1, 5, 2, 6).contiguous()
    w = w.reshape(cout // 128, 128, 9, cin // 16, 4, 4)
    return w.permute(0, 2, 3, 4, 1, 5).contiguous()


def bev_conv3x3_num_tiles(B, H, W):
    n = ctypes.c_int(0)
    _lib.call("crb3d_bev_conv3x3_num_tiles", int(B), int(H), int(W), byref(n))
    return int(n.value)


def bev_tile_plan(coords, n_dev, B, H, W, n_levels):
    """Sparse-tile plan of a stack of n_levels 3x3 stride-1 BEV convs whose first input is the dense() of the sparse tensor with
    rows `coords` (n,4) [b,z,y,x] (n_dev: device row count or None): per level the 128-pixel tiles that can differ from the
    level's constant (see csrc/bev_conv_tc.cu). Returns dict(lists (L,T) int32, counts (L,) int32, flags (L,T) uint8 = computed,
    fill_flags (L,T) uint8 = constant tiles somebody reads)."""
    _need_cuda(coords)
    coords = _i32c(coords)
    dev = coords.device
    T = bev_conv3x3_num_tiles(B, H, W)
    lists = torch.empty((n_levels, T), dtype=torch.int32, device=dev)
    counts = torch.empty((n_levels,), dtype=torch.int32, device=dev)
    flags = torch.empty((n_levels, T), dtype=torch.uint8, device=dev)
    fill_flags = torch.empty((n_levels, T), dtype=torch.uint8, device=dev)
    ws = _ws(_ws_bytes("crb3d_bev_tile_plan_workspace_bytes", int(B), int(H), int(W)), dev)
    _lib.call("crb3d_bev_tile_plan", _p(coords), coords.shape[0], _p(n_dev), int(B), int(H), int(W), int(n_levels), _p(lists), _p(counts),
              _p(flags), _p(fill_flags), _p(ws), ws.numel(), _stream(dev))
    return {"lists": lists, "counts": counts, "flags": flags, "fill_flags": fill_flags, "n_tiles": T}


def bev_conv3x3(x_nhwc, wpack, bias, relu=True, out=None, round_out=False, tiles=None):
    """3x3 / stride 1 / pad 1 conv (+bias, ReLU) on the tensor cores. x_nhwc: (B, H, W, C_in) contiguous fp32 CUDA;
    wpack: pack_conv3x3_weight(...). Returns (B, H, W, C_out) contiguous. tiles = (plan, level, fill): compute only the plan's
    active tiles of that level (bev_tile_plan) and write the constant `fill` (C_out floats) into the others."""
    _need_cuda(x_nhwc, wpack)
    assert x_nhwc.dtype == torch.float32 and x_nhwc.is_contiguous() and wpack.is_contiguous()
    B, H, W, cin = x_nhwc.shape
    cout = wpack.shape[0] * 128
    assert wpack.shape[2] * 16 == cin and (wpack.dim() == 7) == (not (CONV_VARIANT & 1))
    if out is None:
        out = torch.empty((B, H, W, cout), dtype=torch.float32, device=x_nhwc.device)
    prof = PROFILE
    timed = prof is not None and prof.get("mode") == "time" and "conv2d" in prof
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(x_nhwc.device))
    flags_arg = int(bool(relu)) | (2 if round_out else 0) | (CONV_VARIANT << 8)
    pixels = B * H * W
    if tiles is None or (CONV_VARIANT & 1):
        _lib.call("crb3d_bev_conv3x3_tf32", _p(x_nhwc), B, H, W, cin, _p(wpack), cout, _p(_f32c(bias)) if bias is not None else None,
                  flags_arg, _p(out), _stream(x_nhwc.device))
    else:
        plan, level, fill = tiles
        assert plan["n_tiles"] == bev_conv3x3_num_tiles(B, H, W) and fill.numel() == cout and fill.dtype == torch.float32
        _lib.call("crb3d_bev_conv3x3_tf32_tiles", _p(x_nhwc), B, H, W, cin, _p(wpack), cout, _p(_f32c(bias)) if bias is not None else None,
                  flags_arg, _p(out), _p(plan["lists"][level]), _p(plan["counts"][level:level + 1]), _p(plan["fill_flags"][level]),
                  _p(fill), _stream(x_nhwc.device))
    if timed:
        e1.record(torch.cuda.current_stream(x_nhwc.device))
        flops = 2.0 * 9 * cin * cout * pixels
        if tiles is not None and not (CONV_VARIANT & 1):       # flops of the tiles that went through the tensor cores: read after the pass
            cnt, per_tile = tiles[0]["counts"][tiles[1]], 2.0 * 9 * cin * cout * 128
            flops = lambda: float(cnt.item()) * per_tile
        prof["conv2d"].append((e0, e1, flops))
    return out


def pack_conv_gemm_weight(weight):
    """(C_out, C_in, k, k) conv weight -> [C_out][k*k][C_in] TF32-rounded, the layout of crb3d_bev_conv_gemm_tf32."""
    cout, cin, kh, kw = weight.shape
    return round_tf32(weight.detach().float().permute(0, 2, 3, 1).reshape(cout, kh * kw * cin).contiguous())


def bev_conv_gemm(x_nhwc, w2, bias, ksize, stride, pad, relu=True, round_out=False):
    """k x k conv (+bias, ReLU) as an implicit GEMM over strided TMA boxes (csrc/bev_gemm_tc.cu, CONV mode).
    x_nhwc: (B, H, W, C_in) contiguous fp32 CUDA; w2: pack_conv_gemm_weight(...). Returns (B, H_out, W_out, C_out)."""
    _need_cuda(x_nhwc, w2)
    assert x_nhwc.dtype == torch.float32 and x_nhwc.is_contiguous() and w2.is_contiguous()
    B, H, W, cin = x_nhwc.shape
    cout = w2.shape[0]
    assert w2.shape[1] == ksize * ksize * cin
    ho, wo = (H + 2 * pad - ksize) // stride + 1, (W + 2 * pad - ksize) // stride + 1
    out = torch.empty((B, ho, wo, cout), dtype=torch.float32, device=x_nhwc.device)
    prof = PROFILE
    timed = prof is not None and prof.get("mode") == "time" and "conv2d" in prof
    if timed:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(torch.cuda.current_stream(x_nhwc.device))
    _lib.call("crb3d_bev_conv_gemm_tf32", _p(x_nhwc), B, H, W, cin, _p(w2), cout, ksize, stride, pad,
              _p(_f32c(bias)) if bias is not None else None, int(bool(relu)) | (2 if round_out else 0) | (4 if GEMM_CLUSTERS else 0) | (0 if GEMM_PAIRS else 8), _p(out),
              _stream(x_nhwc.device))
    if timed:
        e1.record(torch.cuda.current_stream(x_nhwc.device))
        prof["conv2d"].append((e0, e1, 2.0 * ksize * ksize * cin * cout * B * ho * wo))
    return out


# ----------------------------------------------------------------------------------------------- data path / selection
def mask_collate_points(points, frame_offsets, pc_range, xcol=0, sync=True):
    """mask_points_by_range + collate_batch's batch-index column for a whole batch on the device (csrc/data_prims.cu).
    points (N, C) raw clouds back to back, frame_offsets (B+1) int32. Returns (out (M, 1+C) [b, cols...], out_offsets (B+1));
    sync=False keeps out capacity-sized (N rows) and the counts on the device."""
    _need_cuda(points, frame_offsets)
    points, frame_offsets = _f32c(points), _i32c(frame_offsets)
    n, stride = points.shape
    B = frame_offsets.numel() - 1
    out = torch.empty((max(n, 1), stride + 1), dtype=torch.float32, device=points.device)
    out_off = torch.empty((B + 1,), dtype=torch.int32, device=points.device)
    ws = _ws(_ws_bytes("crb3d_mask_collate_points_workspace_bytes", n), points.device)
    r = (ctypes.c_float * 4)(float(pc_range[0]), float(pc_range[1]), float(pc_range[3]), float(pc_range[4]))
    _lib.call("crb3d_mask_collate_points", _p(points), n, stride, int(xcol), _p(frame_offsets), B, r, _p(out), _p(out_off), _p(ws),
              ws.numel(), _stream(points.device))
    if not sync:
        return out, out_off
    return out[: int(out_off[-1].item())], out_off


def furthest_first(X, min_dist, n_pick):
    """Greedy k-centre picks (coreset_sampling.py:31-52) on the device: X (m, d) f32, min_dist (m) f32 start distances (updated
    in place). Returns a LongTensor (n_pick,) on X's device; no host synchronisation."""
    _need_cuda(X, min_dist)
    X = _f32c(X)
    _check(min_dist=(min_dist, torch.float32))
    m, d = X.shape
    out = torch.empty((n_pick,), dtype=torch.int64, device=X.device)
    ws = _ws(_ws_bytes("crb3d_furthest_first_workspace_bytes", m), X.device)
    _lib.call("crb3d_furthest_first", _p(X), m, d, _p(min_dist), int(n_pick), _p(out), _p(ws), ws.numel(), _stream(X.device))
    return out


# ----------------------------------------------------------------------------------------------- remaining op families
def voxel_query(M, R1, R2, R3, nsample, radius, z_range, y_range, x_range, new_xyz, xyz, new_coords, point_indices, idx):
    """voxel_query_wrapper_stack (pointnet2_stack/src/voxel_query.cpp:28-45): same argument order."""
    _need_cuda(new_xyz, xyz, new_coords, point_indices, idx)
    _check(new_xyz=(new_xyz, torch.float32), xyz=(xyz, torch.float32), new_coords=(new_coords, torch.int32),
           point_indices=(point_indices, torch.int32), idx=(idx, torch.int32))
    _lib.call("crb3d_voxel_query_stack", int(M), int(R1), int(R2), int(R3), int(nsample), float(radius), int(z_range), int(y_range),
              int(x_range), _p(new_xyz), _p(xyz), _p(new_coords), _p(point_indices), _p(idx), _stream(idx.device))


def ball_query_batch(b, n, m, radius, nsample, new_xyz, xyz, idx):
    """ball_query_wrapper_fast (pointnet2_batch/src/ball_query.cpp). idx must come in zero-filled, as the reference's does."""
    _need_cuda(new_xyz, xyz, idx)
    _check(new_xyz=(new_xyz, torch.float32), xyz=(xyz, torch.float32), idx=(idx, torch.int32))
    _lib.call("crb3d_ball_query_batch", int(b), int(n), int(m), float(radius), int(nsample), _p(new_xyz), _p(xyz), _p(idx),
              _stream(idx.device))


def group_points_batch(b, c, n, npoints, nsample, points, idx, out):
    _need_cuda(points, idx, out)
    _check(points=(points, torch.float32), idx=(idx, torch.int32), out=(out, torch.float32))
    _lib.call("crb3d_group_points_batch", int(b), int(c), int(n), int(npoints), int(nsample), _p(points), _p(idx), _p(out),
              _stream(out.device))


def group_points_grad_batch(b, c, n, npoints, nsample, grad_out, idx, grad_points):
    _need_cuda(grad_out, idx, grad_points)
    _check(grad_out=(grad_out, torch.float32), idx=(idx, torch.int32), grad_points=(grad_points, torch.float32))
    _lib.call("crb3d_group_points_grad_batch", int(b), int(c), int(n), int(npoints), int(nsample), _p(grad_out), _p(idx),
              _p(grad_points), _stream(grad_out.device))


def three_nn_batch(b, n, m, unknown, known, dist2, idx):
    _need_cuda(unknown, known, dist2, idx)
    _check(unknown=(unknown, torch.float32), known=(known, torch.float32), dist2=(dist2, torch.float32), idx=(idx, torch.int32))
    _lib.call("crb3d_three_nn_batch", int(b), int(n), int(m), _p(unknown), _p(known), _p(dist2), _p(idx), _stream(idx.device))


def three_interpolate_batch(b, c, m, n, points, idx, weight, out):
    _need_cuda(points, idx, weight, out)
    _check(points=(points, torch.float32), idx=(idx, torch.int32), weight=(weight, torch.float32), out=(out, torch.float32))
    _lib.call("crb3d_three_interpolate_batch", int(b), int(c), int(m), int(n), _p(points), _p(idx), _p(weight), _p(out),
              _stream(out.device))


def three_interpolate_grad_batch(b, c, n, m, grad_out, idx, weight, grad_points):
    _need_cuda(grad_out, idx, weight, grad_points)
    _check(grad_out=(grad_out, torch.float32), idx=(idx, torch.int32), weight=(weight, torch.float32),
           grad_points=(grad_points, torch.float32))
    _lib.call("crb3d_three_interpolate_grad_batch", int(b), int(c), int(n), int(m), _p(grad_out), _p(idx), _p(weight),
              _p(grad_points), _stream(grad_out.device))


def roipoint_pool3d_forward(xyz, boxes3d, pts_feature, pooled, empty_flag):
    """roipool3d_gpu (roipoint_pool3d/src/roipoint_pool3d.cpp:23-46): sizes come from the tensors, as there."""
    _need_cuda(xyz, boxes3d, pts_feature, pooled, empty_flag)
    _check(xyz=(xyz, torch.float32), boxes3d=(boxes3d, torch.float32), pts_feature=(pts_feature, torch.float32),
           pooled=(pooled, torch.float32), empty_flag=(empty_flag, torch.int32))
    B, N = xyz.shape[0], xyz.shape[1]
    M, C, S = boxes3d.shape[1], pts_feature.shape[2], pooled.shape[2]
    ws = _ws(_ws_bytes("crb3d_roipoint_pool3d_workspace_bytes", B, M, S), xyz.device)
    _lib.call("crb3d_roipoint_pool3d_forward", B, N, M, C, S, _p(xyz), _p(boxes3d), _p(pts_feature), _p(pooled), _p(empty_flag),
              _p(ws), ws.numel(), _stream(xyz.device))


# ----------------------------------------------------------------------------------------------- vector pool (PV-RCNN++)
def query_stacked_local_neighbor_idxs(support_xyz, xyz_batch_cnt, new_xyz, new_xyz_batch_cnt, stack_neighbor_idxs, start_len, cumsum,
                                      avg_length_of_neighbor_idxs, max_neighbour_distance, nsample, neighbor_type):
    """query_stacked_local_neighbor_idxs_wrapper_stack (vector_pool.cpp:35-75), same argument order."""
    _need_cuda(support_xyz, new_xyz, start_len, cumsum)
    _check(support_xyz=(support_xyz, torch.float32), xyz_batch_cnt=(xyz_batch_cnt, torch.int32), new_xyz=(new_xyz, torch.float32),
           new_xyz_batch_cnt=(new_xyz_batch_cnt, torch.int32), stack_neighbor_idxs=(stack_neighbor_idxs, torch.int32),
           start_len=(start_len, torch.int32), cumsum=(cumsum, torch.int32))
    M, dev = new_xyz.shape[0], new_xyz.device
    ws = _ws(_ws_bytes("crb3d_query_stacked_local_neighbor_idxs_workspace_bytes", M), dev)
    _lib.call("crb3d_query_stacked_local_neighbor_idxs", _p(support_xyz), _p(xyz_batch_cnt), _p(new_xyz), _p(new_xyz_batch_cnt),
              int(xyz_batch_cnt.shape[0]), M, _p(stack_neighbor_idxs), _p(start_len), _p(cumsum), int(avg_length_of_neighbor_idxs),
              float(max_neighbour_distance), int(nsample), int(neighbor_type), _p(ws), ws.numel(), _stream(dev))


def query_three_nn_by_stacked_local_idxs(support_xyz, new_xyz, new_xyz_grid_centers, new_xyz_grid_idxs, new_xyz_grid_dist2,
                                         stack_neighbor_idxs, start_len, M, num_total_grids):
    """query_three_nn_by_stacked_local_idxs_wrapper_stack (vector_pool.cpp:78-113), same argument order."""
    _need_cuda(support_xyz, new_xyz_grid_centers, new_xyz_grid_idxs, new_xyz_grid_dist2, start_len)
    _check(support_xyz=(support_xyz, torch.float32), new_xyz_grid_centers=(new_xyz_grid_centers, torch.float32),
           new_xyz_grid_idxs=(new_xyz_grid_idxs, torch.int32), new_xyz_grid_dist2=(new_xyz_grid_dist2, torch.float32),
           stack_neighbor_idxs=(stack_neighbor_idxs, torch.int32), start_len=(start_len, torch.int32))
    if stack_neighbor_idxs.numel() == 0:          # no neighbour anywhere: a valid (never dereferenced) address
        stack_neighbor_idxs = torch.zeros(1, dtype=torch.int32, device=support_xyz.device)
    _lib.call("crb3d_query_three_nn_by_stacked_local_idxs", _p(support_xyz), _p(new_xyz_grid_centers), _p(new_xyz_grid_idxs),
              _p(new_xyz_grid_dist2), _p(stack_neighbor_idxs), _p(start_len), int(M), int(num_total_grids), _stream(support_xyz.device))


def vector_pool(support_xyz, xyz_batch_cnt, support_features, new_xyz, new_xyz_batch_cnt, new_features, new_local_xyz,
                point_cnt_of_grid, grouped_idxs, num_grid_x, num_grid_y, num_grid_z, max_neighbour_distance, use_xyz,
                num_max_sum_points, nsample, neighbor_type, pooling_type):
    """vector_pool_wrapper_stack (vector_pool.cpp:116-170), same argument order; returns the cumulative number of grouped points
    as a host int like the reference (one device -> host read)."""
    _need_cuda(support_xyz, support_features, new_xyz, new_features)
    _check(support_xyz=(support_xyz, torch.float32), xyz_batch_cnt=(xyz_batch_cnt, torch.int32),
           support_features=(support_features, torch.float32), new_xyz=(new_xyz, torch.float32),
           new_xyz_batch_cnt=(new_xyz_batch_cnt, torch.int32), new_features=(new_features, torch.float32),
           new_local_xyz=(new_local_xyz, torch.float32), point_cnt_of_grid=(point_cnt_of_grid, torch.int32),
           grouped_idxs=(grouped_idxs, torch.int32))
    dev = support_xyz.device
    cum = torch.zeros(1, dtype=torch.int32, device=dev)
    _lib.call("crb3d_vector_pool_stack", _p(support_xyz), _p(support_features), _p(xyz_batch_cnt), _p(new_xyz), _p(new_xyz_batch_cnt),
              int(xyz_batch_cnt.shape[0]), int(new_xyz.shape[0]), int(support_features.shape[1]), int(new_features.shape[1]),
              int(num_grid_x), int(num_grid_y), int(num_grid_z), float(max_neighbour_distance), int(use_xyz), int(num_max_sum_points),
              int(nsample), int(neighbor_type), int(pooling_type), _p(new_features), _p(new_local_xyz), _p(point_cnt_of_grid),
              _p(grouped_idxs), _p(cum), _stream(dev))
    return int(cum.item())


def vector_pool_grad(grad_new_features, point_cnt_of_grid, grouped_idxs, grad_support_features):
    """vector_pool_grad_wrapper_stack (vector_pool.cpp:173-204), same argument order."""
    _need_cuda(grad_new_features, point_cnt_of_grid, grouped_idxs, grad_support_features)
    _check(grad_new_features=(grad_new_features, torch.float32), point_cnt_of_grid=(point_cnt_of_grid, torch.int32),
           grouped_idxs=(grouped_idxs, torch.int32), grad_support_features=(grad_support_features, torch.float32))
    _lib.call("crb3d_vector_pool_grad_stack", _p(grad_new_features), _p(point_cnt_of_grid), _p(grouped_idxs), _p(grad_support_features),
              int(grad_new_features.shape[1]), int(grad_support_features.shape[1]), int(point_cnt_of_grid.shape[1]),
              int(grouped_idxs.shape[0]), _stream(grad_new_features.device))


# ----------------------------------------------------------------------------------------------- PV-RCNN fused layers
def sa_group_mlp_maxpool(xyz, xyz_cnt, feat, new_xyz, new_cnt, idx, widths, packed, out):
    """One scale of StackSAModuleMSG fused (csrc/sa_mlp.cu): group + MLP + max-pool. `out` is a (M, stride) VIEW whose first
    widths[-1] columns receive the pooled features (column slice of the concatenated multi-scale output)."""
    _need_cuda(xyz, new_xyz, idx, packed, out)
    xyz, new_xyz = _f32c(xyz), _f32c(new_xyz)
    xyz_cnt, new_cnt, idx = _i32c(xyz_cnt), _i32c(new_cnt), _i32c(idx)
    C = 0 if feat is None else feat.shape[1]
    if feat is not None:
        feat = _f32c(feat)
    M, ns = idx.shape
    assert out.dtype == torch.float32 and out.stride(1) == 1 and out.shape[0] == M and out.shape[1] >= widths[-1]
    w = (ctypes.c_int * len(widths))(*[int(x) for x in widths])
    _lib.call("crb3d_sa_group_mlp_maxpool", int(xyz_cnt.numel()), _p(xyz), _p(xyz_cnt), _p(feat), C, _p(new_xyz), _p(new_cnt), M,
              _p(idx), ns, len(widths) - 1, w, _p(_f32c(packed)), _p(out), out.stride(0), _stream(xyz.device))
    return out


def fc_gemm(a, weight, scale=None, shift=None, relu=False):
    """relu?((a (M,K) @ weight (N,K)^T) * scale + shift): split-K tcgen05 GEMM for few rows and a long K (csrc/fc_gemm_tc.cu)."""
    _need_cuda(a, weight)
    assert a.dtype == torch.float32 and a.dim() == 2 and a.stride(1) == 1
    weight = _f32c(weight)
    M, K = a.shape
    N = weight.shape[0]
    out = torch.empty((M, N), dtype=torch.float32, device=a.device)
    ws = _ws(_ws_bytes("crb3d_fc_gemm_workspace_bytes", M, N, K), a.device)
    _lib.call("crb3d_fc_gemm_tf32", _p(a), M, K, a.stride(0), _p(weight), N, _p(_f32c(scale)) if scale is not None else None,
              _p(_f32c(shift)) if shift is not None else None, int(bool(relu)), _p(out), _p(ws), ws.numel(), _stream(a.device))
    return out


# ----------------------------------------------------------------------------------------------- dense
def sparse_to_dense(feat, coords, batch_size, spatial_shape, channels_last_bev=False, out=None, n_dev=None):
    """(B,C,D,H,W) dense tensor (reference .dense()); channels_last_bev=True returns (B,H,W,C*D) memory whose
    .permute(0,3,1,2) equals dense.view(B, C*D, H, W). `out`: optional preallocated destination (zero-filled here)."""
    _need_cuda(feat, coords)
    feat = _f32c(feat)
    coords = _i32c(coords)
    n, C = feat.shape
    D, H, W = [int(x) for x in spatial_shape]
    shape = (batch_size, H, W, C * D) if channels_last_bev else (batch_size, C, D, H, W)
    if out is not None:
        assert tuple(out.shape) == shape and out.is_contiguous() and out.dtype == torch.float32
    dense = out if out is not None else torch.empty(shape, dtype=torch.float32, device=feat.device)
    _lib.call("crb3d_sparse_to_dense", _p(feat), _p(coords), n, C, batch_size, D, H, W, int(channels_last_bev), 1,
              _p(dense), _p(n_dev), _stream(feat.device))
    return dense


def dense_to_sparse(dense, coords, C, spatial_shape, channels_last_bev=False):
    _need_cuda(dense, coords)
    dense = _f32c(dense)
    coords = _i32c(coords)
    n = coords.shape[0]
    D, H, W = [int(x) for x in spatial_shape]
    B = dense.shape[0]
    feat = torch.empty((n, C), dtype=torch.float32, device=dense.device)
    _lib.call("crb3d_dense_to_sparse", _p(dense), _p(coords), n, C, B, D, H, W, int(channels_last_bev), _p(feat),
              _stream(dense.device))
    return feat


# ----------------------------------------------------------------------------------------------- iou3d / nms
def boxes_overlap_bev(boxes_a, boxes_b, out=None):
    _need_cuda(boxes_a, boxes_b)
    a, b = _f32c(boxes_a[:, :7]), _f32c(boxes_b[:, :7])
    if out is None:
        out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _lib.call("crb3d_boxes_overlap_bev", _p(a), a.shape[0], _p(b), b.shape[0], _p(out), _stream(a.device))
    return out


def boxes_iou_bev(boxes_a, boxes_b, out=None):
    _need_cuda(boxes_a, boxes_b)
    a, b = _f32c(boxes_a[:, :7]), _f32c(boxes_b[:, :7])
    if out is None:
        out = torch.zeros((a.shape[0], b.shape[0]), dtype=torch.float32, device=a.device)
    _lib.call("crb3d_boxes_iou_bev", _p(a), a.shape[0], _p(b), b.shape[0], _p(out), _stream(a.device))
    return out


def nms_sorted(boxes_sorted, thresh, rotated=True, max_keep=0):
    """Greedy NMS over boxes already sorted by descending score. Returns (keep int64 device (n,), num_keep device int)
    without any host synchronisation."""
    _need_cuda(boxes_sorted)
    b = _f32c(boxes_sorted[:, :7])
    n = b.shape[0]
    keep = torch.empty((max(n, 1),), dtype=torch.int64, device=b.device)
    num = torch.zeros((1,), dtype=torch.int32, device=b.device)
    ws = _ws(_ws_bytes("crb3d_nms_workspace_bytes", n), b.device)
    _lib.call("crb3d_nms", _p(b), n, float(thresh), int(bool(rotated)), int(max_keep), _p(keep), _p(num), _p(ws),
              ws.numel(), _stream(b.device))
    return keep, num


def nms_mask(boxes_sorted, thresh, rotated=True):
    _need_cuda(boxes_sorted)
    b = _f32c(boxes_sorted[:, :7])
    n = b.shape[0]
    cb = (n + 63) // 64
    mask = torch.zeros((n, cb), dtype=torch.int64, device=b.device)
    ws = _ws(_ws_bytes("crb3d_nms_workspace_bytes", n), b.device)
    _lib.call("crb3d_nms_mask", _p(b), n, float(thresh), int(bool(rotated)), _p(mask), _p(ws), ws.numel(), _stream(b.device))
    return mask


def boxes_iou_bev_cpu(boxes_a, boxes_b, out):
    a = boxes_a.float().contiguous()
    b = boxes_b.float().contiguous()
    assert not a.is_cuda and not b.is_cuda and not out.is_cuda and out.dtype == torch.float32 and out.is_contiguous()
    _lib.call("crb3d_boxes_iou_bev_cpu", _p(a), a.shape[0], _p(b), b.shape[0], _p(out))
    return out


# ----------------------------------------------------------------------------------------------- roiaware
def points_in_boxes(boxes, pts, out=None):
    """boxes (B,T,7), pts (B,M,3) -> (B,M) int32 index of the first containing box or -1."""
    _need_cuda(boxes, pts)
    boxes, pts = _f32c(boxes), _f32c(pts)
    B, T, _ = boxes.shape
    M = pts.shape[1]
    if out is None:
        out = torch.full((B, M), -1, dtype=torch.int32, device=pts.device)
    _lib.call("crb3d_points_in_boxes", _p(boxes), _p(pts), B, T, M, _p(out), _stream(pts.device))
    return out


def points_in_boxes_stack(pts, pt_off, boxes, box_off, max_pts_per_frame, want_density=True):
    """Stacked frames. pts (N,S>=3) xyz first; boxes (T,7). Returns (idx (N,), counts (T,), density (T,) | None)."""
    _need_cuda(pts, pt_off, boxes, box_off)
    pts, boxes = _f32c(pts), _f32c(boxes)
    pt_off, box_off = _i32c(pt_off), _i32c(box_off)
    N, S = pts.shape
    T = boxes.shape[0]
    B = pt_off.numel() - 1
    idx = torch.empty((N,), dtype=torch.int32, device=pts.device)
    counts = torch.empty((max(T, 1),), dtype=torch.int32, device=pts.device)
    dens = torch.empty((max(T, 1),), dtype=torch.float32, device=pts.device) if want_density else None
    _lib.call("crb3d_points_in_boxes_stack", _p(pts), S, _p(pt_off), int(max_pts_per_frame), _p(boxes), _p(box_off), B,
              N, T, _p(idx), _p(counts), _p(dens), _stream(pts.device))
    return idx, counts[:T], (dens[:T] if dens is not None else None)


def points_in_boxes_cpu(boxes, pts, out):
    b = boxes.float().contiguous()
    p = pts.float().contiguous()
    assert not b.is_cuda and not p.is_cuda and out.dtype == torch.int32 and out.is_contiguous()
    _lib.call("crb3d_points_in_boxes_cpu", _p(b), b.shape[0], _p(p), p.shape[0], _p(out))
    return out


def roiaware_pool3d_forward(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled, pool_method):
    _need_cuda(rois, pts, pts_feature, argmax, pts_idx_of_voxels, pooled)
    rois, pts, pts_feature = _f32c(rois), _f32c(pts), _f32c(pts_feature)
    _check(argmax=(argmax, torch.int32), pts_idx_of_voxels=(pts_idx_of_voxels, torch.int32), pooled=(pooled, torch.float32))
    n_boxes, ox, oy, oz, C = pooled.shape
    _lib.call("crb3d_roiaware_pool3d_forward", _p(rois), _p(pts), _p(pts_feature), n_boxes, pts.shape[0], C,
              pts_idx_of_voxels.shape[4], ox, oy, oz, _p(argmax), _p(pts_idx_of_voxels), _p(pooled), int(pool_method),
              _stream(pts.device))


def roiaware_pool3d_backward(pts_idx_of_voxels, argmax, grad_out, grad_in, pool_method):
    _need_cuda(pts_idx_of_voxels, argmax, grad_out, grad_in)
    _check(pts_idx_of_voxels=(pts_idx_of_voxels, torch.int32), argmax=(argmax, torch.int32), grad_in=(grad_in, torch.float32))
    n_boxes, ox, oy, oz, C = grad_out.shape
    _lib.call("crb3d_roiaware_pool3d_backward", _p(pts_idx_of_voxels), _p(argmax), _p(_f32c(grad_out)), _p(grad_in),
              n_boxes, ox, oy, oz, C, pts_idx_of_voxels.shape[4], int(pool_method), _stream(grad_in.device))


# ----------------------------------------------------------------------------------------------- pointnet2 (stack)
def ball_query(B, M, radius, nsample, new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx, max_queries_per_frame=0):
    _need_cuda(new_xyz, new_xyz_batch_cnt, xyz, xyz_batch_cnt, idx)
    _check(new_xyz=(new_xyz, torch.float32), new_xyz_batch_cnt=(new_xyz_batch_cnt, torch.int32), xyz=(xyz, torch.float32), xyz_batch_cnt=(xyz_batch_cnt, torch.int32), idx=(idx, torch.int32))
    _check_len(new_xyz_batch_cnt=(new_xyz_batch_cnt, B), xyz_batch_cnt=(xyz_batch_cnt, B))
    _lib.call("crb3d_ball_query_stack", B, M, float(radius), int(nsample), _p(new_xyz), _p(new_xyz_batch_cnt), _p(xyz),
              _p(xyz_batch_cnt), _p(idx), int(max_queries_per_frame), _stream(idx.device))


def group_points(B, M, C, nsample, features, features_batch_cnt, idx, idx_batch_cnt, out):
    _need_cuda(features, features_batch_cnt, idx, idx_batch_cnt, out)
    _check(features=(features, torch.float32), features_batch_cnt=(features_batch_cnt, torch.int32), idx=(idx, torch.int32), idx_batch_cnt=(idx_batch_cnt, torch.int32), out=(out, torch.float32))
    _check_len(features_batch_cnt=(features_batch_cnt, B), idx_batch_cnt=(idx_batch_cnt, B))
    _lib.call("crb3d_group_points_stack", B, M, C, nsample, _p(features), _p(features_batch_cnt), _p(idx),
              _p(idx_batch_cnt), _p(out), _stream(out.device))


def group_points_grad(B, M, C, N, nsample, grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features):
    _need_cuda(grad_out, idx, idx_batch_cnt, features_batch_cnt, grad_features)
    _check(grad_out=(grad_out, torch.float32), idx=(idx, torch.int32), idx_batch_cnt=(idx_batch_cnt, torch.int32), features_batch_cnt=(features_batch_cnt, torch.int32), grad_features=(grad_features, torch.float32))
    _check_len(features_batch_cnt=(features_batch_cnt, B), idx_batch_cnt=(idx_batch_cnt, B))
    _lib.call("crb3d_group_points_grad_stack", B, M, C, N, nsample, _p(grad_out), _p(idx), _p(idx_batch_cnt),
              _p(features_batch_cnt), _p(grad_features), _stream(grad_out.device))


def farthest_point_sampling(b, n, m, points, temp, idx):
    _need_cuda(points, temp, idx)
    _check(points=(points, torch.float32), temp=(temp, torch.float32), idx=(idx, torch.int32))
    _lib.call("crb3d_farthest_point_sampling", b, n, m, _p(points), _p(temp), _p(idx), _stream(idx.device))


def stack_farthest_point_sampling(points, temp, xyz_batch_cnt, idx, num_sampled_points, n_max=None):
    _need_cuda(points, temp, xyz_batch_cnt, idx, num_sampled_points)
    _check(points=(points, torch.float32), temp=(temp, torch.float32), xyz_batch_cnt=(xyz_batch_cnt, torch.int32), idx=(idx, torch.int32), num_sampled_points=(num_sampled_points, torch.int32))
    B = xyz_batch_cnt.numel()
    if n_max is None:
        n_max = int(xyz_batch_cnt.max().item())
    _lib.call("crb3d_stack_farthest_point_sampling", B, int(n_max), _p(points), _p(temp), _p(xyz_batch_cnt), _p(idx),
              _p(num_sampled_points), _stream(idx.device))


def three_nn(B, N, M, unknown, unknown_batch_cnt, known, known_batch_cnt, dist2, idx):
    _need_cuda(unknown, unknown_batch_cnt, known, known_batch_cnt, dist2, idx)
    _check(unknown=(unknown, torch.float32), unknown_batch_cnt=(unknown_batch_cnt, torch.int32), known=(known, torch.float32), known_batch_cnt=(known_batch_cnt, torch.int32), dist2=(dist2, torch.float32), idx=(idx, torch.int32))
    _check_len(unknown_batch_cnt=(unknown_batch_cnt, B), known_batch_cnt=(known_batch_cnt, B))
    _lib.call("crb3d_three_nn_stack", B, N, M, _p(unknown), _p(unknown_batch_cnt), _p(known), _p(known_batch_cnt),
              _p(dist2), _p(idx), _stream(idx.device))


def three_interpolate(N, C, features, idx, weight, out):
    _need_cuda(features, idx, weight, out)
    _check(features=(features, torch.float32), idx=(idx, torch.int32), weight=(weight, torch.float32), out=(out, torch.float32))
    _lib.call("crb3d_three_interpolate_stack", N, C, _p(features), _p(idx), _p(weight), _p(out), _stream(out.device))


def three_interpolate_grad(N, C, grad_out, idx, weight, grad_features):
    _need_cuda(grad_out, idx, weight, grad_features)
    _check(grad_out=(grad_out, torch.float32), idx=(idx, torch.int32), weight=(weight, torch.float32), grad_features=(grad_features, torch.float32))
    _lib.call("crb3d_three_interpolate_grad_stack", N, C, _p(grad_out), _p(idx), _p(weight), _p(grad_features),
              _stream(grad_out.device))


# ----------------------------------------------------------------------------------------------- CRB scoring
def label_entropy(labels, box_off, num_class, want_counts=False):
    """Per-frame CRB stage-1 entropy of predicted labels (1-based), stacked with box_off (B+1)."""
    _need_cuda(labels, box_off)
    labels, box_off = _i32c(labels), _i32c(box_off)
    B = box_off.numel() - 1
    ent = torch.empty((B,), dtype=torch.float32, device=labels.device)
    cc = torch.empty((B, num_class), dtype=torch.int32, device=labels.device) if want_counts else None
    _lib.call("crb3d_label_entropy", _p(labels), _p(box_off), B, num_class, _p(ent), _p(cc), _stream(labels.device))
    return (ent, cc) if want_counts else ent


def pairwise_sqdist(X):
    _need_cuda(X)
    X = _f32c(X)
    n, d = X.shape
    D = torch.empty((n, n), dtype=torch.float64, device=X.device)
    _lib.call("crb3d_pairwise_sqdist_f64", _p(X), n, d, _p(D), _stream(X.device))
    return D


def kde_greedy(dens, labels, cand_off, n_class, axis, prior_n, bandwidth, n_select):
    """CRB stage-3 greedy selection. Returns (order (n_select,) int32 device, picked_score (n_select,) f64 device)."""
    _need_cuda(dens, labels, cand_off, axis, prior_n)
    dens, labels, cand_off = _f32c(dens), _i32c(labels), _i32c(cand_off)
    axis = axis.double().contiguous()
    prior_n = prior_n.double().contiguous()
    n_cand = cand_off.numel() - 1
    dev = dens.device
    order = torch.full((n_select,), -1, dtype=torch.int32, device=dev)
    ps = torch.zeros((n_select,), dtype=torch.float64, device=dev)
    ws = _ws(_ws_bytes("crb3d_kde_greedy_workspace_bytes", n_cand, n_class), dev)
    _lib.call("crb3d_kde_greedy", _p(dens), _p(labels), _p(cand_off), n_cand, n_class, _p(axis), _p(prior_n),
              float(bandwidth), int(n_select), _p(order), _p(ps), _p(ws), ws.numel(), _stream(dev))
    return order, ps


# ----------------------------------------------------------------------------------------------- batched scoring path
def nms_batched(boxes, counts, thresh, rotated=True, max_keep=0):
    """boxes (B, n_max, 7) (each frame sorted by descending score), counts (B,) int32 device.
    Returns keep (B, max_keep or n_max) int64 and num_keep (B,) int32 - no host sync."""
    _need_cuda(boxes, counts)
    boxes = _f32c(boxes)
    counts = _i32c(counts)
    B, n_max, _ = boxes.shape
    stride = max_keep if max_keep > 0 else n_max
    keep = torch.zeros((B, max(stride, 1)), dtype=torch.int64, device=boxes.device)
    num = torch.zeros((B,), dtype=torch.int32, device=boxes.device)
    ws = _ws(_ws_bytes("crb3d_nms_batched_workspace_bytes", B, n_max), boxes.device)
    _lib.call("crb3d_nms_batched", _p(boxes), _p(counts), B, n_max, float(thresh), int(bool(rotated)), int(max_keep),
              _p(keep), max(stride, 1), _p(num), _p(ws), ws.numel(), _stream(boxes.device))
    return keep, num


def points_in_boxes_ranges(pts, pt_begin, pt_end, max_pts_per_frame, boxes, box_begin, box_end, want_density=True):
    """Like points_in_boxes_stack but with explicit per-frame [begin, end) row ranges into `pts` and `boxes`
    (padded box tensors). counts/density are indexed like boxes' rows."""
    _need_cuda(pts, pt_begin, pt_end, boxes, box_begin, box_end)
    pts = _f32c(pts)
    boxes2 = _f32c(boxes.reshape(-1, boxes.shape[-1])[:, :7])
    N, S = pts.shape
    slots = boxes2.shape[0]
    B = pt_begin.numel()
    idx = torch.full((N,), -1, dtype=torch.int32, device=pts.device)
    counts = torch.zeros((max(slots, 1),), dtype=torch.int32, device=pts.device)
    dens = torch.zeros((max(slots, 1),), dtype=torch.float32, device=pts.device) if want_density else None
    _lib.call("crb3d_points_in_boxes_ranges", _p(pts), S, _p(_i32c(pt_begin)), _p(_i32c(pt_end)), int(max_pts_per_frame),
              _p(boxes2), _p(_i32c(box_begin)), _p(_i32c(box_end)), B, slots, _p(idx), _p(counts), _p(dens),
              _stream(pts.device))
    return idx, counts[:slots], (dens[:slots] if dens is not None else None)


def label_entropy_ranges(labels, box_begin, box_end, num_class):
    _need_cuda(labels, box_begin, box_end)
    labels = _i32c(labels.reshape(-1))
    B = box_begin.numel()
    ent = torch.empty((B,), dtype=torch.float32, device=labels.device)
    _lib.call("crb3d_label_entropy_ranges", _p(labels), _p(_i32c(box_begin)), _p(_i32c(box_end)), B, num_class, _p(ent),
              None, _stream(labels.device))
    return ent
