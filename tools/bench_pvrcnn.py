#!/usr/bin/env python
"""BASELINE.json configs[2]: PV-RCNN KITTI-synthetic on 1 x B200. The detector is the REFERENCE's own PVRCNN class (built from
its pv_rcnn_active_crb.yaml, imported from /root/reference or the staging baseline/_ref by tests/ref_env.py) running over the
crb3d drop-in; measured (a) as it is - every compiled op already a kernel of this library, the module code the reference's -
and (b) after crb3d.pvrcnn.accelerate (fused SA layers, split-K tensor-core FC, MC rounds sharing the first FC, tensor-core
BEV plan). Eval forward + post_processing with MC dropout on, batch 4, CUDA events, 3 warm-up + 10 timed iterations."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "tests"), os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

import ref_env  # noqa: E402
from test_gpu_pvrcnn import _calibrate_bn, _reference_pvrcnn  # noqa: E402

cuda = torch.device("cuda:0")
B = int(sys.argv[1]) if len(sys.argv) > 1 else 4
model, cfg, _batch = _reference_pvrcnn(cuda)
bd, frames, pts, offs_t = _batch(cuda, B, cfg.DATA_CONFIG)
_calibrate_bn(model, bd)
for m in model.modules():
    if m.__class__.__name__.startswith("Dropout"):
        m.train()


def rate(n=10):
    with torch.no_grad():
        for _ in range(3):
            model(dict(bd))
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        for _ in range(n):
            model(dict(bd))
        e1.record()
        torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n


def stage_times():
    """ms per module of the detector's module_list (+ post_processing), CUDA events, one pass after warm-up."""
    names = [type(m).__name__ for m in model.module_list] + ["post_processing"]
    ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
    with torch.no_grad():
        d = dict(bd)
        ev[0].record()
        for i, m in enumerate(model.module_list):
            d = m(d)
            ev[i + 1].record()
        model.post_processing(d)
        ev[-1].record()
    torch.cuda.synchronize()
    return {n: round(ev[i].elapsed_time(ev[i + 1]), 3) for i, n in enumerate(names)}


res = {"batch": B, "points_per_frame": int(np.mean([len(f) for f in frames]))}
ms = rate()
res["reference_modules_over_dropin"] = {"ms_per_batch": ms, "frames_per_s": B / ms * 1e3, "stages_ms": stage_times()}
from crb3d import pvrcnn  # noqa: E402
pvrcnn.accelerate(model)
ms = rate()
res["accelerated"] = {"ms_per_batch": ms, "frames_per_s": B / ms * 1e3, "stages_ms": stage_times()}
print(json.dumps(res))
