#!/usr/bin/env python
"""Round-2 micro-benchmark of the BEV block-2 kernels (run on the GPU box): the CTA-pair 3x3 kernel with 1- and 2-tile work
items and the implicit-GEMM conv (strided TMA boxes) against cuDNN TF32, L2 flushed between launches."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

from crb3d import ops  # noqa: E402

flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")


def timeit(fn, n=20):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(n):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


def case(B, H, W, cin, cout, stride):
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    b = torch.randn(cout, generator=g).cuda()
    ho, wo = (H + 2 - 3) // stride + 1, (W + 2 - 3) // stride + 1
    fl = 2.0 * 9 * cin * cout * B * ho * wo
    torch.backends.cudnn.allow_tf32 = True
    xc, wc = x.permute(0, 3, 1, 2), w.contiguous(memory_format=torch.channels_last)
    t_dnn = timeit(lambda: torch.cudnn_convolution_relu(xc, wc, b, (stride, stride), (1, 1), (1, 1), 1))
    w2 = ops.pack_conv_gemm_weight(w)
    t_gemm = timeit(lambda: ops.bev_conv_gemm(x, w2, b, 3, stride, 1, True))
    res = "conv3x3 s%d B%d %dx%d %d->%d: cuDNN %.1f us (%.0f TF/s) | implicit GEMM %.1f us (%.0f TF/s)" % (
        stride, B, H, W, cin, cout, t_dnn, fl / t_dnn / 1e6, t_gemm, fl / t_gemm / 1e6)
    if stride == 1:
        wp = ops.pack_conv3x3_weight(w, split=True)
        for name, var in (("pair auto", 0), ("pair auto, release.cluster arrive", 16), ("pair 1-tile items", 4), ("pair 2-tile items", 8)):
            ops.CONV_VARIANT = var
            t = timeit(lambda: ops.bev_conv3x3(x, wp, b, True))
            res += " | %s %.1f us (%.0f TF/s)" % (name, t, fl / t / 1e6)
        ops.CONV_VARIANT = 0
    print(res, flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "b16":         # the bench's batch: item-size choice of the pair kernel
        case(16, 100, 88, 256, 256, 1)
        case(16, 200, 176, 128, 128, 1)
        case(16, 200, 176, 256, 128, 1)
        sys.exit(0)
    case(4, 200, 176, 128, 256, 2)
    case(4, 100, 88, 256, 256, 1)
    case(4, 200, 176, 128, 128, 1)
    case(4, 200, 176, 256, 128, 1)
    case(8, 100, 88, 256, 256, 1)
    case(2, 188, 188, 128, 256, 2)
    case(2, 94, 94, 256, 256, 1)
