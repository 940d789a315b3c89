#!/usr/bin/env python
"""Times crb3d_nms_batched on score-sorted clustered boxes (B frames x n candidates), GPU box only."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT, os.path.join(ROOT, "tests")):
    sys.path.insert(0, p)
import numpy as np  # noqa: E402
import torch  # noqa: E402

from crb3d import ops  # noqa: E402
from util import rand_boxes  # noqa: E402


def case(B, n, thr, max_keep, extent, cluster):
    rng = np.random.default_rng(n + B)
    boxes = torch.from_numpy(np.stack([rand_boxes(rng, n, extent, cluster) for _ in range(B)])).cuda()
    counts = torch.full((B,), n, dtype=torch.int32, device="cuda")
    for _ in range(3):
        keep, num = ops.nms_batched(boxes, counts, thr, rotated=True, max_keep=max_keep)
    ts = []
    for _ in range(10):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        keep, num = ops.nms_batched(boxes, counts, thr, rotated=True, max_keep=max_keep)
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1) * 1e3)
    ts.sort()
    print("nms B%d n%d thr%.2f max_keep%d extent%g cluster%d: %.1f us, kept %s" %
          (B, n, thr, max_keep, extent, cluster, ts[len(ts) // 2], num.cpu().tolist()), flush=True)


if __name__ == "__main__":
    case(4, 4096, 0.01, 500, 40, True)
    case(4, 4096, 0.01, 500, 40, False)
    case(4, 4096, 0.7, 500, 40, True)
    case(1, 9000, 0.8, 512, 40, True)
    case(1, 1024, 0.7, 128, 40, True)
    case(4, 300, 0.1, 0, 20, True)
