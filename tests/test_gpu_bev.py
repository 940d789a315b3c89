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
                                        (1, 32, 256, True), (129, 64, 80, True)])
def test_bev_gemm_plain(cuda, M, K, N, relu):
    from crb3d import ops
    _fp32_reference_mode()
    g = torch.Generator(device="cpu").manual_seed(M + K + N)
    a = torch.randn(M, K, generator=g).to(cuda)
    w = (torch.randn(N, K, generator=g) / np.sqrt(K)).to(cuda)
    b = torch.randn(N, generator=g).to(cuda)
    out = torch.full((M, N + 8), -7.0, device=cuda)
    ops.bev_gemm(a, w, b, relu, [(out[:, 4:], 0, N, N + 8)])
    ref = a.double() @ w.double().t() + b.double()
    if relu:
        ref = ref.clamp_min(0)
    _check(out[:, 4:4 + N], ref, "gemm")
    assert float(out[:, :4].min()) == -7.0 and float(out[:, 4 + N:].max()) == -7.0   # nothing outside the segment is touched


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
