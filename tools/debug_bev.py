import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np, torch
import bench
from crb3d import scorer, second, ops
dev = torch.device("cuda:0")
model = bench.build_model(dev)
model.prepare_inference(fold_bev_bn=True, spconv_tf32=True)
frames, batches = bench.make_batches(4)
ps = scorer.PoolScorer(model, dev, 4)
r = ps.to_device(ps.stage_host(batches[0]))
with torch.no_grad():
    geom = model.geometry(r[0], r[1], 4)
    bd = model.backbone_3d(dict(batch_size=4, **geom))
    for k, t in bd["multi_scale_3d_features"].items():
        print(k, "rows", t.features.shape[0], "absmax %.3e rms %.3e" % (float(t.features.abs().max()), float(t.features.pow(2).mean().sqrt())))
    enc = bd["encoded_spconv_tensor"]
    print("enc absmax %.3e" % float(enc.features.abs().max()))
    bd = model.map_to_bev_module(bd)
    x = bd["spatial_features"]
    print("spatial absmax %.3e nonzero frac %.4f" % (float(x.abs().max()), float((x != 0).float().mean())))
    bb = model.backbone_2d
    xh = x.permute(0, 2, 3, 1).contiguous()
    for bi, (layers, de, gemm) in enumerate(bb._plan):
        for li, (w, b, stride, pad, wpack) in enumerate(layers):
            own = wpack is not None and bb._tc_conv_pays(4, xh.shape[1], xh.shape[2], w.shape[0], w.shape[1])
            ref = torch.cudnn_convolution_relu(xh.permute(0, 3, 1, 2), w, b, stride, pad, (1, 1), 1).permute(0, 2, 3, 1).contiguous()
            if own:
                xh2 = ops.bev_conv3x3(xh, wpack, b, True, round_out=True)
                print("block", bi, "layer", li, "own: absmax %.3e vs cudnn %.3e, max diff %.3e" % (float(xh2.abs().max()), float(ref.abs().max()), float((xh2 - ref).abs().max())))
                xh = xh2
            else:
                print("block", bi, "layer", li, "cudnn: absmax %.3e" % float(ref.abs().max()))
                xh = ref
    out = bb(dict(spatial_features=x))["spatial_features_2d"]
    print("x2d absmax %.3e rms %.3e" % (float(out.abs().max()), float(out.pow(2).mean().sqrt())))
