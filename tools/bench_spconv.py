#!/usr/bin/env python
"""Per-layer timing of the sparse-conv kernels on the real SECOND rulebooks of a synthetic KITTI batch.
  python tools/bench_spconv.py [--batch 4] [--reps 30]
Prints, per backbone layer: rows, pairs, algorithmic bytes, us and GB/s for the tcgen05 and SIMT kernels."""
import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "crb-active-3ddet_b200"))
import numpy as np
import torch

from crb3d import ops, second, synth

ap = argparse.ArgumentParser()
ap.add_argument("--batch", type=int, default=4)
ap.add_argument("--reps", type=int, default=30)
ap.add_argument("--only", type=str, default="")
ap.add_argument("--no-simt", action="store_true")
args = ap.parse_args()
dev = torch.device("cuda:0")
torch.manual_seed(0)
model = second.SECONDNet().eval().to_device(dev)
frames = [synth.make_frame(i) for i in range(args.batch)]
offs = torch.from_numpy(np.cumsum([0] + [len(f) for f in frames]).astype(np.int32)).to(dev)
pts = torch.from_numpy(np.concatenate(frames)).to(dev)
geom = model.geometry(pts, offs, args.batch)
books = geom["rulebooks"]
layers = [("conv_input", "subm1", 4, 16), ("conv1", "subm1", 16, 16), ("conv2.0", "spconv2", 16, 32), ("conv2.1", "subm2", 32, 32),
          ("conv3.0", "spconv3", 32, 64), ("conv3.1", "subm3", 64, 64), ("conv4.0", "spconv4", 64, 64), ("conv4.1", "subm4", 64, 64),
          ("conv_out", "spconv_down2", 64, 128)]
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
print("%-10s %8s %8s %9s %7s | %9s %8s | %9s %8s | %9s" % ("layer", "n_in", "n_out", "pairs", "MB", "grouped us", "GB/s", "1/stage us", "GB/s", "simt us"))
for name, key, cin, cout in layers:
    if args.only and args.only != name:
        continue
    d = books[key]
    nbr = d.nbr
    K, n_out = nbr.shape
    n_in = d.indices.shape[0]
    pairs = int((nbr >= 0).sum())
    feat = torch.randn(n_in, cin, device=dev)
    w = torch.randn(cout, K, cin, device=dev) * 0.05
    alg = 4 * (pairs * cin + n_out * cout + K * cin * cout) + 8 * pairs
    res = []
    ref_out = None
    for mode in ("grouped", "legacy", "simt"):
        tf32 = mode != "simt"
        ops.SPCONV_GROUPED = mode == "grouped"
        if mode == "simt" and args.no_simt:
            res.append(float("nan"))
            continue
        for _ in range(3):
            o = ops.spconv_forward(feat, nbr, w, tf32=tf32)
        if mode == "grouped":
            ref_out = o
        elif mode == "legacy":
            err = float((o - ref_out).abs().max() / o.abs().max().clamp_min(1e-9))
            assert err < 1e-5, (name, err)      # same products, same order of the offsets: identical up to fp32 accumulation order
        ts = []
        for _ in range(args.reps):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.spconv_forward(feat, nbr, w, tf32=tf32)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1) * 1e3)
        res.append(float(np.median(ts)))
    print("%-10s %8d %8d %9d %7.1f | %9.1f %8.1f | %9.1f %8.1f | %9.1f" % (name, n_in, n_out, pairs, alg / 1e6, res[0], alg / res[0] / 1e3,
                                                                       res[1], alg / res[1] / 1e3, res[2]))
