#!/usr/bin/env python
"""A/B timing of the BEV GEMM variants at the bench shapes: the CTA-pair kernel (csrc/bev_gemm_pair.cu, streamed / resident weights)
against the single-CTA kernel and its opt-in cluster-multicast variants (csrc/bev_gemm_tc.cu) on deblock 2, deblock 1 and the
stride-2 3x3 conv; every variant is checked bit for bit against the single-CTA kernel.
  python tools/bench_gemm_cluster.py [--batch 16]"""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crb-active-3ddet_b200"))
import torch

from crb3d import ops

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=16)
ap.add_argument("--reps", type=int, default=20)
args = ap.parse_args()
dev = torch.device("cuda:0")
B = args.batch
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)


def timeit(fn):
    for _ in range(3):
        fn()
    ts = []
    for _ in range(args.reps):
        flush.fill_(1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    return ts[len(ts) // 2]


g = torch.Generator(device=dev).manual_seed(0)
# deblock 2: (B,100,88,256) -> ConvTranspose2d(256, 256, 2, 2) into channels [256, 512) of the (B,200,176,512) map
x2 = torch.randn(B, 100, 88, 256, device=dev, generator=g)
w2 = ops.round_tf32(torch.randn(4 * 256, 256, device=dev, generator=g) / 16)
b2 = torch.randn(256, device=dev, generator=g)
cat = torch.zeros(B, 200, 176, 512, device=dev)
outs = {}
ops.GEMM_CLUSTERS = False
for pairs in (True, False):      # CTA pairs (csrc/bev_gemm_pair.cu) against the single-CTA kernel
    ops.GEMM_PAIRS = pairs
    cat.zero_()
    fn = lambda: ops.bev_gemm(x2.view(-1, 256), w2, b2, True, [(cat[..., 256:], 0, 256, 512)], n_sub=4, up=2, in_hw=(100, 88), round_out=True)
    t = timeit(fn)
    outs["pairs%d" % pairs] = cat.clone()
    print("deblock2 B%d pairs=%s: %.1f us (%.0f TFLOP/s)" % (B, pairs, t, 2.0 * B * 100 * 88 * 256 * 1024 / t / 1e6))
ops.GEMM_PAIRS, ops.GEMM_PAIR_STREAM = True, True
cat.zero_()
t = timeit(fn)
print("deblock2 B%d pairs, streamed weights: %.1f us, equal=%s" % (B, t, torch.equal(cat, outs["pairs1"])))
ops.GEMM_PAIR_STREAM = False
d = (outs["pairs1"] - outs["pairs0"]).abs().max().item()
print("pairs vs single: max abs diff %.3g (max |value| %.3g), equal=%s" % (d, outs["pairs0"].abs().max().item(), torch.equal(outs["pairs1"], outs["pairs0"])))
ref = torch.relu(torch.nn.functional.conv_transpose2d(x2.permute(0, 3, 1, 2).double(), w2.view(2, 2, 256, 256).permute(3, 2, 0, 1).double(),
                                                      b2.double(), stride=2)).permute(0, 2, 3, 1)
err = (outs["pairs1"][..., 256:].double() - ref).pow(2).mean().sqrt().item() / ref.pow(2).mean().sqrt().item()
print("pairs vs fp64 conv_transpose2d: rel RMS %.3g" % err)
ops.GEMM_PAIRS = False
for mode in (True, False):
    ops.GEMM_CLUSTERS = mode
    fn = lambda: ops.bev_gemm(x2.view(-1, 256), w2, b2, True, [(cat[..., 256:], 0, 256, 512)], n_sub=4, up=2, in_hw=(100, 88), round_out=True)
    t = timeit(fn)
    outs[mode] = cat.clone()
    flops = 2.0 * B * 100 * 88 * 256 * 1024
    print("deblock2 B%d clusters=%s: %.1f us (%.0f TFLOP/s, %.2f TB/s algorithmic)" % (B, mode, t, flops / t / 1e6,
                                                                                   (x2.numel() + B * 200 * 176 * 256) * 4 / t / 1e6))
assert torch.equal(outs[True], outs[False])
# deblock 1: (B,200,176,128) -> ConvTranspose2d(128, 256, 1, 1) = a plain GEMM into channels [0, 256)
x1 = torch.randn(B, 200, 176, 128, device=dev, generator=g)
w1 = ops.round_tf32(torch.randn(256, 128, device=dev, generator=g) / 11)
ops.GEMM_CLUSTERS, ops.GEMM_PAIRS = False, True
r1 = {}
for shortk, stream in ((False, False), (True, False), (True, True)):
    ops.GEMM_PAIR_SHORTK, ops.GEMM_PAIR_STREAM = shortk, stream
    cat.zero_()
    fn = lambda: ops.bev_gemm(x1.view(-1, 128), w1, b2, True, [(cat, 0, 256, 512)], round_out=True)
    t = timeit(fn)
    r1[(shortk, stream)] = cat.clone()
    print("deblock1 B%d pair kernel=%s streamed=%s: %.1f us (%.2f TB/s)" % (B, shortk, stream, t, (x1.numel() + B * 200 * 176 * 256) * 4 / t / 1e6))
print("deblock1 equal:", torch.equal(r1[(False, False)], r1[(True, False)]), torch.equal(r1[(False, False)], r1[(True, True)]))
ops.GEMM_PAIR_SHORTK, ops.GEMM_PAIR_STREAM = False, True
# stride-2 conv: (B,200,176,128) -> (B,100,88,256)
x = torch.randn(B, 200, 176, 128, device=dev, generator=g)
w = ops.round_tf32(torch.randn(256, 128, 3, 3, device=dev, generator=g) / 30)
wp = ops.pack_conv_gemm_weight(w)
bb = torch.randn(256, device=dev, generator=g)
res = {}
ops.GEMM_CLUSTERS, ops.GEMM_PAIRS = False, True
fn = lambda: ops.bev_conv_gemm(x, wp, bb, 3, 2, 1, True, round_out=True)
t = timeit(fn)
res["pairs"] = fn()
print("conv3x3 s2 B%d pairs: %.1f us (%.0f TFLOP/s)" % (B, t, 2.0 * B * 100 * 88 * 9 * 128 * 256 / t / 1e6))
ops.GEMM_PAIRS = False
for mode in (True, False):
    ops.GEMM_CLUSTERS = mode
    fn = lambda: ops.bev_conv_gemm(x, wp, bb, 3, 2, 1, True, round_out=True)
    t = timeit(fn)
    res[mode] = fn()
    flops = 2.0 * B * 100 * 88 * 9 * 128 * 256
    print("conv3x3 s2 B%d clusters=%s: %.1f us (%.0f TFLOP/s)" % (B, mode, t, flops / t / 1e6))
assert torch.equal(res[True], res[False])
print("conv pairs vs single: equal=%s, max abs diff %.3g" % (torch.equal(res["pairs"], res[False]), (res["pairs"] - res[False]).abs().max().item()))
