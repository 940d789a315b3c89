#!/usr/bin/env python
"""How many anchors pass SCORE_THRESH per frame in the bench configuration (debug aid)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from crb3d import scorer, second
dev = torch.device("cuda:0")
model = bench.build_model(dev)
model.prepare_inference(fold_bev_bn=True, spconv_tf32=True)
frames, batches = bench.make_batches(4)
ps = scorer.PoolScorer(model, dev, 4)
res = [ps.to_device(ps.stage_host(b)) for b in batches]
second.calibrate_batchnorm(model, res[0][0], res[0][1], 4)
second.calibrate_head_bias(model, res[0][0], res[0][1], 4, target_fraction=0.004)
with torch.no_grad():
    for i, r in enumerate(res):
        bd = model.forward_features(r[0], r[1], 4)
        sc = torch.sigmoid(bd["cls_preds"].max(-1).values)
        print("batch", i, "anchors >= 0.1 per frame:", (sc >= 0.1).sum(1).tolist(), "distinct score values:", [int(torch.unique(sc[b][sc[b] >= 0.1]).numel()) for b in range(4)])
    bd = model.forward_features(res[0][0], res[0][1], 4)
    lg = bd["cls_preds"].reshape(-1, 18)
    print("post-calibration logits: mean per channel", [round(float(x), 2) for x in lg.mean(0)], "std", [round(float(x), 2) for x in lg.std(0)])
    print("fraction of entries above logit(0.1):", float((lg > -2.197).float().mean()))
    w = model.dense_head.conv_cls.weight
    print("conv_cls weight absmax", float(w.abs().max()), "bias", [round(float(x), 2) for x in model.dense_head.conv_cls.bias])
    plan = model.dense_head._plan
    print("plan weight absmax", float(plan[0][:18].abs().max()), "plan bias", [round(float(x), 2) for x in plan[1][:18]])
    x2d = bd["spatial_features_2d"]
    ref = torch.nn.functional.conv2d(x2d, model.dense_head.conv_cls.weight, model.dense_head.conv_cls.bias).permute(0, 2, 3, 1).reshape(-1, 18)
    print("module-path logits: mean", [round(float(x), 2) for x in ref.mean(0)][:6], "max abs diff vs plan", float((ref - lg).abs().max()))
