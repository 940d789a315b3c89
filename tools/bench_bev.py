#!/usr/bin/env python
"""Micro-benchmark + parity probe for the tcgen05 BEV kernels (run on the GPU box):
  python tools/bench_bev.py            -> one line per shape: max/rms error vs fp32 torch, us own kernel, us cuDNN TF32
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    sys.path.insert(0, p)
import torch  # noqa: E402

from crb3d import ops  # noqa: E402


ONCE = len(sys.argv) > 1 and sys.argv[1] == "once"     # one launch per case (for ncu captures)


def timeit(fn, n=20):
    if ONCE:
        fn()
        torch.cuda.synchronize()
        return 1.0
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
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


def conv_case(B, H, W, cin, cout, check=True):
    g = torch.Generator().manual_seed(cin + cout + H)
    x = torch.randn(B, H, W, cin, generator=g).cuda()
    w = (torch.randn(cout, cin, 3, 3, generator=g) / (3 * cin ** 0.5)).cuda()
    b = torch.randn(cout, generator=g).cuda()
    wp = ops.pack_conv3x3_weight(w)
    out = ops.bev_conv3x3(x, wp, b, True)
    torch.cuda.synchronize()
    msg = ""
    if check and not ONCE:
        torch.backends.cudnn.allow_tf32 = False
        ref = torch.relu(torch.nn.functional.conv2d(x.permute(0, 3, 1, 2), w, b, padding=1)).permute(0, 2, 3, 1)
        err = (out - ref).abs()
        rms = ref.pow(2).mean().sqrt()
        msg = "max_err/rms %.2e rms_err/rms %.2e" % (float(err.max() / rms), float(err.pow(2).mean().sqrt() / rms))
    torch.backends.cudnn.allow_tf32 = True
    xc = x.permute(0, 3, 1, 2)
    wc = w.contiguous(memory_format=torch.channels_last)
    t_own = timeit(lambda: ops.bev_conv3x3(x, wp, b, True, out=out))
    t_dnn = 1.0 if ONCE else timeit(lambda: torch.cudnn_convolution_relu(xc, wc, b, (1, 1), (1, 1), (1, 1), 1))
    fl = 2.0 * 9 * cin * cout * B * H * W
    print("conv3x3 B%d %dx%d %d->%d: %s | own %.1f us (%.0f TF/s) cudnn %.1f us (%.0f TF/s)" %
          (B, H, W, cin, cout, msg, t_own, fl / t_own / 1e6, t_dnn, fl / t_dnn / 1e6), flush=True)


def gemm_case(name, M, K, N, n_sub=1, up=0, in_hw=(0, 0), segs_w=None, ctot=None):
    g = torch.Generator().manual_seed(M % 1000 + K + N)
    a = torch.randn(M, K, generator=g).cuda()
    w = (torch.randn(n_sub * N, K, generator=g) / K ** 0.5).cuda()
    b = torch.randn(N, generator=g).cuda()
    if segs_w is None:
        ctot = ctot or N
        out = torch.empty((M * n_sub, ctot), device="cuda")
        segs = [(out, 0, N, ctot)]
        out_bytes = M * n_sub * N * 4
    else:
        outs = [torch.empty((M, n), device="cuda") for n in segs_w]
        segs, c0 = [], 0
        for o, n in zip(outs, segs_w):
            segs.append((o, c0, n, n))
            c0 += n
        out_bytes = M * sum(segs_w) * 4
    t = timeit(lambda: ops.bev_gemm(a, w, b, True, segs, n_sub=n_sub, up=up, in_hw=in_hw))
    byts = M * K * 4 + out_bytes + w.numel() * 4
    print("gemm %s M%d K%d N%d x%d: own %.1f us, %.0f GB/s algorithmic, %.0f TF/s" %
          (name, M, K, N, n_sub, t, byts / t / 1e3, 2.0 * M * K * N * n_sub / t / 1e6), flush=True)


if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "variants":      # kernel experiments of bev_conv_tc.cu (results of 2, 3 are wrong by design)
        for v in (0, 1):
            ops.CONV_VARIANT = v
            print("variant", v, flush=True)
            conv_case(1, 16, 16, 16, 128)
            conv_case(1, 37, 29, 64, 256)
            conv_case(4, 200, 176, 128, 128)
            conv_case(4, 200, 176, 256, 128)
            conv_case(4, 100, 88, 256, 256)
            conv_case(8, 200, 176, 128, 128, check=False)
        sys.exit(0)
    if len(sys.argv) > 1 and sys.argv[1] == "trace":         # per-CTA phase breakdown of the conv kernels (experiment bit 1)
        import ctypes
        import numpy as np
        from crb3d import _lib
        lib = _lib.load()
        lib.crb3d_bev_conv3x3_trace.argtypes = [ctypes.c_void_p, ctypes.c_int]
        g = torch.Generator().manual_seed(1)
        x = torch.randn(4, 200, 176, 128, generator=g).cuda()
        w = (torch.randn(128, 128, 3, 3, generator=g) / 30).cuda()
        b = torch.randn(128, generator=g).cuda()
        out = torch.empty(4, 200, 176, 128, device="cuda")
        flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
        buf = np.zeros((296, 16), dtype=np.int64)
        for v in (2, 3):
            ops.CONV_VARIANT = v
            wp = ops.pack_conv3x3_weight(w)
            n = 148 if v == 2 else 275
            for _ in range(3):
                ops.bev_conv3x3(x, wp, b, True, out=out)
            torch.cuda.synchronize()
            lib.crb3d_bev_conv3x3_trace(buf.ctypes.data, n)
            flush.fill_(1)
            ops.bev_conv3x3(x, wp, b, True, out=out)
            torch.cuda.synchronize()
            lib.crb3d_bev_conv3x3_trace(buf.ctypes.data, n)
            d = buf[:n]
            t0 = d[:, 0].min()
            print("variant %d (%s): kernel span %.1f us" % (v, "CTA pair" if v == 2 else "single CTA", (max(d[:, 5].max(), d[:, 10].max()) - t0) / 1e3))
            if v == 2:
                ld = d[0::2]            # leaders
                mhz = np.median(ld[:, 13] / np.maximum(ld[:, 3] - ld[:, 2], 1)) * 1e3
                print("  leaders: setup %.2f | first data %.2f | mma loop %.2f (%.0f clk at ~%.0f MHz, %d..%d items) | last epilogue ends +%.2f | exit +%.2f us"
                      % (np.median(ld[:, 1] - ld[:, 0]) / 1e3, np.median(ld[:, 2] - ld[:, 1]) / 1e3, np.median(ld[:, 3] - ld[:, 2]) / 1e3,
                         np.median(ld[:, 13]), mhz, ld[:, 14].min(), ld[:, 14].max(), np.median(ld[:, 10] - ld[:, 3]) / 1e3,
                         np.median(ld[:, 5] - ld[:, 3]) / 1e3))
                for items in sorted(set(ld[:, 14])):
                    sel = ld[ld[:, 14] == items]
                    print("    %d leaders with %d items: mma loop %.2f us = %.0f clk/item; waits (clk) full_a %d full_b %d acc_empty %d; B producer on empty_b %d; epilogue warp on acc_full %d"
                          % (len(sel), items, np.median(sel[:, 3] - sel[:, 2]) / 1e3, np.median(sel[:, 13]) / items, np.median(sel[:, 9]),
                             np.median(sel[:, 8]), np.median(sel[:, 11]), np.median(sel[:, 7]), np.median(sel[:, 12])))
            else:
                first = d[:, 0] - t0 < 5000          # CTAs of the first wave
                for name, sel in (("wave 1", first), ("wave 2", ~first)):
                    e = d[sel]
                    print("  %s: start +%.1f us | setup %.2f | first data %.2f | main loop %.2f | drain %.2f | epilogue %.2f us"
                          % (name, np.median(e[:, 0] - t0) / 1e3, np.median(e[:, 1] - e[:, 0]) / 1e3, np.median(e[:, 2] - e[:, 1]) / 1e3,
                             np.median(e[:, 3] - e[:, 2]) / 1e3, np.median(e[:, 4] - e[:, 3]) / 1e3, np.median(e[:, 10] - e[:, 4]) / 1e3))
                    print("    waits (clk): producer on empty_b %d, mma on full_b %d, mma on full_a %d" %
                          (np.median(e[:, 7]), np.median(e[:, 8]), np.median(e[:, 9])))
        sys.exit(0)
    gemm_case("deconv1-contig", 4 * 200 * 176, 128, 256, ctot=256)
    gemm_case("deconv1", 4 * 200 * 176, 128, 256, ctot=512)
    gemm_case("deconv2", 4 * 100 * 88, 256, 256, n_sub=4, up=2, in_hw=(100, 88), ctot=512)
    gemm_case("heads", 4 * 200 * 176, 512, 80, segs_w=(18, 42, 12))
    conv_case(1, 16, 16, 16, 128)
    conv_case(2, 24, 40, 32, 128)
    conv_case(1, 37, 29, 64, 256)
    conv_case(4, 200, 176, 128, 128)
    conv_case(4, 200, 176, 256, 128)
    conv_case(4, 100, 88, 256, 256)
    conv_case(8, 100, 88, 256, 256, check=False)
    conv_case(2, 188, 188, 128, 128, check=False)
