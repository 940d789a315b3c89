"""-m gpu parity: the tcgen05 BEV GEMM / conv kernels (csrc/bev_gemm_tc.cu, bev_conv_tc.cu) against a plain PyTorch
fp32 reference of the same op (TF32 switched off for the reference). Tolerance: TF32 inputs (10-bit mantissa) with fp32
accumulation (the tensor core truncates the fp32 operands to TF32) -> |err| <= 8e-3 * RMS(reference) per element and
<= 1e-3 * RMS(reference) in RMS."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _fp32_reference_mode():
    torch.backends.cuda.matmul.allow_tf32 = False
    torch.backends.cudnn.allow_tf32 = False


def _check(out, ref, what):
    ref = ref.double()
    err = (out.double() - ref).abs()
    rms = float(ref.pow(2).mean().sqrt())
    assert float(err.max()) <= 8e-3 * rms, "%s: max err %.3e vs rms %.3e" % (what, float(err.max()), rms)
    assert float(err.pow(2).mean().sqrt()) <= 1e-3 * rms, "%s: rms err %.3e vs rms %.3e" % (what, float(err.pow(2).mean().sqrt()), rms)


@pytest.mark.parametrize("M,K,N,relu", [(128 * 7 + 5, 128, 256, True), (1000, 256, 128, False), (35200, 512, 80, False),
                                        (1, 32, 256, True), (129, 64, 80, True), (128 * 300 + 17, 256, 256, True),
                                        (20000, 512, 256, False), (5000, 384, 128, True)])
def test_bev_gemm_plain(cuda, M, K, N, relu):
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    if N == 80:      # N = 80 is the dense-rows placement (row_stride == width)
        out = torch.full((M, N), -7.0, device=cuda)
        ops.bev_gemm(a, w, b, relu, [(out, 0, N, N)])
        got = out
    else:            # full rows inside a wider buffer: columns [4, 4+N) of an (M, N+8) tensor
        out = torch.full((M, N + 8), -7.0, device=cuda)
        ops.bev_gemm(a, w, b, relu, [(out[:, 4:], 0, N, N + 8)])
        got = out[:, 4:4 + N]
        assert float(out[:, :4].min()) == -7.0 and float(out[:, 4 + N:].max()) == -7.0   # nothing outside the segment is touched
    ref = a.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    _check(got, ref, "gemm")


def test_bev_gemm_three_segments(cuda):
    """The anchor-head layout: one N=80 GEMM (72 live columns) scattered to (M,18), (M,42), (M,12)."""
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(5)
    M, K = 176 * 50 + 3, 512
    a = torch.randn(M, K, generator=g).to(cuda)
    w = torch.zeros(80, K)
    w[:72] = torch.randn(72, K, generator=g) / np.sqrt(K)
    b = torch.zeros(80)
    b[:72] = torch.randn(72, generator=g)
    w, b = w.to(cuda), b.to(cuda)
    outs = [torch.empty((M, n), device=cuda) for n in (18, 42, 12)]
    ops.bev_gemm(a, w, b, False, [(outs[0], 0, 18, 18), (outs[1], 18, 42, 42), (outs[2], 60, 12, 12)])
    ref = a.double() @ w.double().t() + b.double()
    _check(torch.cat(outs, 1), ref[:, :72], "head gemm")


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 12, 11, 256, 256), (1, 100, 88, 256, 256), (3, 5, 7, 64, 128)])
def test_bev_deconv2x2_into_concat_slice(cuda, B, H, W, cin, cout):
    """ConvTranspose2d(kernel = stride = 2) + bias + ReLU written into channels [c0, c0+cout) of a wider NHWC map."""
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(B * H + W)
    x = torch.randn(B, cin, H, W, generator=g).to(cuda).contiguous(memory_format=torch.channels_last)
    wt = (torch.randn(cin, cout, 2, 2, generator=g) / np.sqrt(cin)).to(cuda)
    bias = torch.randn(cout, generator=g).to(cuda)
    ctot, c0 = cout + 64, 32
    cat = torch.full((B, 2 * H, 2 * W, ctot), 3.0, device=cuda)
    a = x.permute(0, 2, 3, 1).reshape(B * H * W, cin)
    gw = wt.permute(2, 3, 1, 0).reshape(4 * cout, cin).contiguous()
    ops.bev_gemm(a, gw, bias, True, [(cat.view(-1, ctot)[:, c0:], 0, cout, ctot)], n_sub=4, up=2, in_hw=(H, W))
    ref = torch.relu(torch.nn.functional.conv_transpose2d(x.double(), wt.double(), bias.double(), stride=2))
    _check(cat[..., c0:c0 + cout].permute(0, 3, 1, 2), ref, "deconv")
    assert float(cat[..., :c0].min()) == 3.0 and float(cat[..., c0 + cout:].min()) == 3.0


def test_second_dense_half_matches_torch_modules(cuda):
    """BaseBEVBackbone + AnchorHeadSingle inference plan (folded BN, tcgen05 GEMM deblocks/heads) == the nn.Module path."""
    from crb3d import second
    _fp32_reference_mode()
    torch.manual_seed(1)
    model = second.SECONDNet().eval().to_device(cuda)
    with torch.no_grad():
        for m in model.backbone_2d.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_mean.normal_(0, 0.1)
                m.running_var.uniform_(0.5, 1.5)
                m.weight.uniform_(0.5, 1.5)
                m.bias.normal_(0, 0.1)
        x = torch.randn(2, 256, 200, 176, device=cuda).contiguous(memory_format=torch.channels_last)
        ref = model.dense_head(model.backbone_2d(dict(spatial_features=x)))
        ref = {k: ref[k].clone() for k in ("spatial_features_2d", "cls_preds", "box_preds", "dir_cls_preds")}
        torch.backends.cudnn.allow_tf32 = True     # the plan's 3x3 convs run in TF32, like the reference's default
        model.prepare_inference(fold_bev_bn=True)
        out = model.dense_head(model.backbone_2d(dict(spatial_features=x)))
    for k in ref:
        r = ref[k].double()
        err = (out[k].double() - r).abs()
        rms = float(r.pow(2).mean().sqrt())
        assert tuple(out[k].shape) == tuple(ref[k].shape)
        assert float(err.pow(2).mean().sqrt()) <= 2e-3 * rms, (k, float(err.pow(2).mean().sqrt()), rms)


@pytest.mark.parametrize("B,H,W,cin,cout", [(1, 16, 16, 16, 128), (2, 24, 40, 32, 128), (1, 37, 29, 64, 256), (3, 8, 16, 128, 128),
                                            (2, 200, 176, 128, 128), (1, 100, 88, 256, 256), (5, 9, 7, 16, 128)])
def test_bev_conv3x3_halo_tile(cuda, B, H, W, cin, cout):
    """3x3 / stride 1 / pad 1 conv + bias + ReLU: ragged edges (H, W not multiples of the 8x16 tile), CTAs with fewer
    than four tiles, both tile orientations and C_out split over two CTAs."""
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(H * W + cin)
    x = torch.randn(B, H, W, cin, generator=g).to(cuda)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * np.sqrt(cin))).to(cuda)
    b = torch.randn(cout, generator=g).to(cuda)
    out = ops.bev_conv3x3(x, ops.pack_conv3x3_weight(w), b, True)
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)
    _check(out, ref, "conv3x3")
    lin = ops.bev_conv3x3(x, ops.pack_conv3x3_weight(w), None, False)
    _check(lin, torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), None, padding=1).permute(0, 2, 3, 1), "conv3x3 linear")


@pytest.mark.parametrize("B,H,W,cin,cout", [(2, 24, 40, 32, 128), (1, 37, 29, 64, 256), (5, 9, 7, 16, 128)])
def test_bev_conv3x3_single_cta_kernel(cuda, B, H, W, cin, cout, monkeypatch):
    """The single-CTA predecessor of the CTA-pair kernel (flag bit 8, unsplit weight layout) stays a valid comparison arm."""
    from crb3d import ops
    _fp32_reference_mode()
    monkeypatch.setattr(ops, "CONV_VARIANT", 1)
    g = torch.Generator(device="cpu").manual_seed(H * W + cin + 1)
    x = torch.randn(B, H, W, cin, generator=g).to(cuda)
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * np.sqrt(cin))).to(cuda)
    b = torch.randn(cout, generator=g).to(cuda)
    wp = ops.pack_conv3x3_weight(w)
    assert wp.dim() == 6
    out = ops.bev_conv3x3(x, wp, b, True)
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)
    _check(out, ref, "conv3x3 single-CTA")


@pytest.mark.parametrize("B,A,thr,k,shift", [(4, 211200, 0.1, 4096, 0.0), (2, 5000, 0.5, 4096, 4.0), (3, 30000, 0.2, 512, 0.0),
                                             (1, 100, 0.9, 64, 0.0), (2, 70000, 0.6, 4096, 0.0), (2, 64, 0.0, 4096, 0.0),
                                             (2, 9000, 0.5, 4096, 30.0), (4, 211200, 0.1, 4096, 3.0)])
def test_head_scores_topk(cuda, B, A, thr, k, shift):
    """Fused score pass + candidate top-k == (score >= thr) then sort by (score desc, anchor index asc), cut at k - both
    the few-candidates path and the radix-select path (more than k candidates); duplicates in the scores included."""
    from crb3d import head_ops
    g = torch.Generator(device="cpu").manual_seed(A + k)
    cls = (torch.randn(B, A, 3, generator=g) * 2 - 3 + shift).to(cuda)
    m = (A // 7) * 7
    cls[:, 0:m:7] = cls[:, 1:m:7]                                   # exact score ties between different anchors
    s, l, ts, ti, c = head_ops.anchor_head_scores_topk(cls, 3, B, thr, k)
    s2, l2 = head_ops.anchor_head_scores(cls, 3)
    assert torch.equal(s, s2) and torch.equal(l, l2)
    sc = s.view(B, A).cpu().numpy()
    for b in range(B):
        cand = np.nonzero(sc[b] >= np.float32(thr))[0]
        order = cand[np.lexsort((cand, -sc[b][cand].astype(np.float64)))][:k]
        n = len(order)
        assert int(c[b]) == n
        assert np.array_equal(ti[b, :n].cpu().numpy(), order)
        assert np.array_equal(ts[b, :n].cpu().numpy(), sc[b][order])
        assert float(ts[b, n:].abs().sum()) == 0


def test_dropin_accelerate_bev_backbone(cuda):
    """crb3d.dropin.accelerate_bev_backbone on a module with the reference's BaseBEVBackbone structure (plain nn.Module with
    `blocks` / `deblocks`, NCHW-contiguous input like the reference's HeightCompression output)."""
    from crb3d import dropin, second
    _fp32_reference_mode()
    torch.manual_seed(3)
    ref_like = second.BaseBEVBackbone(second.KITTI_SECOND_CFG, 256).eval().to(cuda)     # same structure / parameter names
    with torch.no_grad():
        for m in ref_like.modules():
            if isinstance(m, torch.nn.BatchNorm2d):
                m.running_var.uniform_(0.5, 1.5)
                m.running_mean.normal_(0, 0.1)
        x = torch.randn(4, 256, 200, 176, device=cuda)                                  # NCHW contiguous
        ref = torch.nn.Module.__call__(ref_like, dict(spatial_features=x))["spatial_features_2d"].clone()
        torch.backends.cudnn.allow_tf32 = True
        dropin.accelerate_bev_backbone(ref_like)
        out = ref_like(dict(spatial_features=x))["spatial_features_2d"]
    assert tuple(out.shape) == tuple(ref.shape) == (4, 512, 200, 176)
    err = (out.double() - ref.double()).pow(2).mean().sqrt()
    assert float(err) <= 2e-3 * float(ref.double().pow(2).mean().sqrt())


@pytest.mark.parametrize("B,H,W,cin,cout,k,stride,pad", [(2, 200, 176, 128, 256, 3, 2, 1), (1, 37, 29, 64, 128, 3, 2, 1),
                                                         (3, 24, 40, 32, 256, 3, 1, 1), (1, 9, 7, 32, 128, 3, 2, 1),
                                                         (2, 16, 16, 96, 128, 1, 1, 0), (1, 100, 88, 256, 256, 3, 1, 1)])
def test_bev_conv_gemm_strided_boxes(cuda, B, H, W, cin, cout, k, stride, pad):
    """k x k conv (stride 1 / 2, zero padding) + bias + ReLU as an implicit GEMM over strided 4-D TMA boxes (the stride-2
    first conv of BEV block 2, base_bev_backbone.py:33-40): odd sizes, both strides, C_out split over two CTA slices."""
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(H * W + cin + stride)
    x = torch.randn(B, H, W, cin, generator=g).to(cuda)
    w = (torch.randn(cout, cin, k, k, generator=g) / (k * np.sqrt(cin))).to(cuda)
    b = torch.randn(cout, generator=g).to(cuda)
    out = ops.bev_conv_gemm(x, ops.pack_conv_gemm_weight(w), b, k, stride, pad, True)
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), stride=stride, padding=pad)).permute(0, 2, 3, 1)
    assert tuple(out.shape) == tuple(ref.shape)
    _check(out, ref, "conv gemm")


@pytest.mark.parametrize("n_occ", [0, 1, 400, 6000])
def test_bev_block1_sparse_tiles_bit_identical(cuda, n_occ):
    """Block 1 of the BEV backbone in sparse-tile mode (only tiles near an occupied cell or the border on the tensor cores, the
    rest filled with each layer's constant) == the dense computation, bit for bit, from an empty map (every interior tile is
    constant, most CTA pairs get no work item) to a map where every tile is active."""
    from crb3d import ops, second
    torch.manual_seed(3)
    net = second.SECONDNet().eval().to_device(cuda)
    bb = net.backbone_2d
    bb.build_inference_plan()
    assert bb._sparse_fills is not None and len(bb._sparse_fills) == 6
    B, H, W, C = 2, 200, 176, 256
    rng = np.random.default_rng(n_occ)
    flat = rng.choice(B * H * W, n_occ, replace=False) if n_occ else np.zeros((0,), np.int64)
    b, y, x = np.unravel_index(flat, (B, H, W))
    if n_occ == 400:                                  # clustered: a few blobs, like objects in a scene
        y = (100 + 20 * rng.standard_normal(n_occ)).clip(0, H - 1).astype(np.int64)
        x = (60 + 15 * rng.standard_normal(n_occ)).clip(0, W - 1).astype(np.int64)
    coords = torch.from_numpy(np.stack([b, np.zeros_like(b), y, x], 1).astype(np.int32)).to(cuda)
    dense = torch.zeros((B, H, W, C), device=cuda)
    dense[coords[:, 0].long(), coords[:, 2].long(), coords[:, 3].long()] = torch.randn(len(coords), C, device=cuda)
    xin = dense.permute(0, 3, 1, 2)
    with torch.no_grad():
        # capacity-sized coordinate buffer with a device-side count, as in the captured step
        pad = torch.cat([coords, torch.full((50, 4), 1, dtype=torch.int32, device=cuda)])
        n_dev = torch.tensor([len(coords)], dtype=torch.int32, device=cuda)
        # constant tiles nobody reads are NOT written: poison the allocator's free blocks so that a missing fill cannot hide
        # behind stale values of an earlier (dense) run
        poison = [torch.full((B, H, W, 128), float("nan"), device=cuda) for _ in range(4)]
        del poison
        got = bb.forward_inference(xin, occupancy=(pad, n_dev))
        assert bool(torch.isfinite(got).all())
        ref = bb.forward_inference(xin)
        plan = ops.bev_tile_plan(pad, n_dev, B, H, W, 6)
    assert torch.equal(got, ref)
    ff = plan["fill_flags"].cpu().numpy()
    assert not (ff & plan["flags"].cpu().numpy()).any() and (ff[-1] == 1 - plan["flags"][-1].cpu().numpy()).all()
    if n_occ == 0:
        assert ff[0].sum() < ff[-1].sum()              # early levels fill only the ring next to the computed tiles
    counts = plan["counts"].cpu().numpy()
    T = plan["n_tiles"]
    assert (np.diff(counts) >= 0).all() and counts[-1] <= T        # the active set only grows with depth
    assert np.array_equal(plan["flags"].sum(1).cpu().numpy(), counts)
    border = 2 * (2 * 25 + 2 * 11 - 4)
    if n_occ == 0:
        assert counts[0] == border                                 # only the border tiles (zero padding in reach)
    if n_occ == 6000:
        assert counts[-1] > 0.9 * T
    for l in range(6):                                             # lists = the flagged tiles in ascending order
        lst = plan["lists"][l, :counts[l]].cpu().numpy()
        assert np.array_equal(lst, np.nonzero(plan["flags"][l].cpu().numpy())[0])


def test_gemm_cluster_multicast_is_bit_identical(cuda, monkeypatch):
    """The opt-in thread-block-cluster variants of the long-K GEMMs (TMA multicast of the shared operand k-blocks, multicast
    tcgen05.commit; csrc/bev_gemm_tc.cu) give exactly the plain launch's results: a 2x2 transposed conv (clusters of 4 weight
    slices), a stride-2 3x3 conv with 256 output channels (2 x 2 clusters) and one with 128 (pairs of tiles, odd tile count)."""
    from crb3d import ops
    g = torch.Generator(device="cpu").manual_seed(5)
    x2 = torch.randn(3, 50, 44, 256, generator=g).to(cuda)
    w2 = ops.round_tf32((torch.randn(4 * 256, 256, generator=g) / 16).to(cuda))
    b2 = torch.randn(256, generator=g).to(cuda)
    x = torch.randn(3, 37, 52, 64, generator=g).to(cuda)
    res = {}
    for mode in (False, True):
        monkeypatch.setattr(ops, "GEMM_CLUSTERS", mode)
        cat = torch.zeros(3, 100, 88, 512, device=cuda)
        ops.bev_gemm(x2.view(-1, 256), w2, b2, True, [(cat[..., 256:], 0, 256, 512)], n_sub=4, up=2, in_hw=(50, 44), round_out=True)
        convs = []
        for cout in (256, 128):
            w = ops.round_tf32((torch.randn(cout, 64, 3, 3, generator=torch.Generator().manual_seed(cout)) / 24).to(cuda))
            convs.append(ops.bev_conv_gemm(x, ops.pack_conv_gemm_weight(w), None, 3, 2, 1, True))
        res[mode] = [cat] + convs
    for a, b in zip(res[False], res[True]):
        assert torch.equal(a, b)
    assert float(res[True][0][..., 256:].abs().max()) > 0


def test_bev_gemm_pair_kernel_identical_to_single_cta(cuda, monkeypatch):
    """The CTA-pair GEMM (csrc/bev_gemm_pair.cu: cta_group::2 M256 x N256 MMAs, weights streamed or resident) gives exactly the
    single-CTA kernel's results on a ragged deblock (6600 rows = 25.8 tile pairs, 2x2 pixel interleave into a channel slice)
    and on the stride-2 3x3 conv in CONV mode (19 x 26 outputs: partial tiles in both directions, odd tile count)."""
    from crb3d import ops
    g = torch.Generator(device="cpu").manual_seed(9)
    x2 = torch.randn(3, 50, 44, 256, generator=g).to(cuda)
    w2 = ops.round_tf32((torch.randn(4 * 256, 256, generator=g) / 16).to(cuda))
    b2 = torch.randn(256, generator=g).to(cuda)
    x = torch.randn(3, 37, 52, 64, generator=g).to(cuda)
    w = ops.round_tf32((torch.randn(256, 64, 3, 3, generator=g) / 24).to(cuda))
    bc = torch.randn(256, generator=g).to(cuda)
    monkeypatch.setattr(ops, "GEMM_CLUSTERS", False)
    res = {}
    for name, pairs, stream in (("single", False, False), ("pair_resident", True, False), ("pair_streamed", True, True)):
        monkeypatch.setattr(ops, "GEMM_PAIRS", pairs)
        monkeypatch.setattr(ops, "GEMM_PAIR_STREAM", stream)
        cat = torch.zeros(3, 100, 88, 512, device=cuda)
        ops.bev_gemm(x2.view(-1, 256), w2, b2, True, [(cat[..., 256:], 0, 256, 512)], n_sub=4, up=2, in_hw=(50, 44), round_out=True)
        conv = ops.bev_conv_gemm(x, ops.pack_conv_gemm_weight(w), bc, 3, 2, 1, True, round_out=True)
        res[name] = (cat, conv)
    for name in ("pair_resident", "pair_streamed"):
        assert torch.equal(res[name][0], res["single"][0]), name
        assert torch.equal(res[name][1], res["single"][1]), name
    assert float(res["single"][0][..., :256].abs().max()) == 0.0 and float(res["single"][0][..., 256:].abs().max()) > 0
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), bc.double(), stride=2, padding=1)).permute(0, 2, 3, 1)
    err = float((res["pair_streamed"][1].double() - ref).pow(2).mean().sqrt() / ref.pow(2).mean().sqrt())
    assert err < 1e-3, err          # TF32 operands (weights pre-rounded, activations truncated by the tensor core), fp32 accumulate


@pytest.mark.parametrize("variant", [4, 8])
def test_bev_conv3x3_pair_item_sizes(cuda, variant, monkeypatch):
    """Both work-item sizes of the CTA-pair kernel (flag bits 10 / 11 force 1-tile / 2-tile items) on the block-2 shape."""
    from crb3d import ops
    _fp32_reference_mode()
    monkeypatch.setattr(ops, "CONV_VARIANT", variant)
    g = torch.Generator(device="cpu").manual_seed(variant)
    x = torch.randn(4, 100, 88, 256, generator=g).to(cuda)
    w = (torch.randn(256, 256, 3, 3, generator=g) / 48).to(cuda)
    b = torch.randn(256, generator=g).to(cuda)
    out = ops.bev_conv3x3(x, ops.pack_conv3x3_weight(w, split=True), b, True)
    ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2).double(), w.double(), b.double(), padding=1)).permute(0, 2, 3, 1)
    _check(out, ref, "conv3x3 pair variant %d" % variant)


def test_bev_inference_plan_has_no_cudnn_layer(cuda):
    """Every conv of the KITTI BEV backbone runs on a kernel of this library in the inference plan (round 1 left block 2 and the
    stride-2 conv to cuDNN): the plan must name a packed weight for each layer and the profiler must see no cudnn/cutlass kernel."""
    from crb3d import second
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(cuda)
    model.prepare_inference(fold_bev_bn=True)
    for layers, _, gemm in model.backbone_2d._plan:
        assert gemm is not None
        for w, b, stride, pad, wpack, w2 in layers:
            assert wpack is not None or w2 is not None, (tuple(w.shape), stride)
    x = torch.randn(4, 256, 200, 176, device=cuda).contiguous(memory_format=torch.channels_last)
    with torch.no_grad():
        model.dense_head(model.backbone_2d(dict(spatial_features=x)))
        torch.cuda.synchronize()
        with torch.profiler.profile(activities=[torch.profiler.ProfilerActivity.CUDA]) as prof:
            model.dense_head(model.backbone_2d(dict(spatial_features=x)))
            torch.cuda.synchronize()
    names = [e.key for e in prof.key_averages()]
    assert names and not [n for n in names if "cudnn" in n.lower() or "cutlass" in n.lower() or "implicit_gemm" in n.lower()], names
