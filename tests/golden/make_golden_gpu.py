"""Generates tests/golden/ref_kernels_*.npz from the REFERENCE's own CUDA kernels (oracle/_ref/libpcdet_ref_kernels.so,
compiled from /root/reference by oracle/build.py). Run on a GPU box:
    python tests/golden/make_golden_gpu.py gpurun_out/golden
then copy the .npz files into tests/golden/. The CPU (-m "not gpu") tests pin oracle/csrc/oracle.c against them and the
-m gpu tests pin the CUDA path against them even when oracle/_ref is absent."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import P, rand_boxes, ref_kernels  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = ref_kernels()
    assert ref is not None, "oracle/_ref not built"
    dev = torch.device("cuda:0")
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    rng = np.random.default_rng(2024)

    # --- iou3d_nms
    a, b = rand_boxes(rng, 96, 12, True), rand_boxes(rng, 80, 12, True)
    ta, tb = cu(a), cu(b)
    ov, iou = torch.zeros((96, 80), device=dev), torch.zeros((96, 80), device=dev)
    ref.ref_boxes_overlap(96, P(ta), 80, P(tb), P(ov)); ref.ref_boxes_iou_bev(96, P(ta), 80, P(tb), P(iou))
    n = 300
    nb = rand_boxes(rng, n, 14, True)
    tnb = cu(nb)
    cb = (n + 63) // 64
    keeps = {}
    for name, thr, fn in (("rot", 0.25, ref.ref_nms_mask), ("normal", 0.5, ref.ref_nms_normal_mask)):
        mask = torch.zeros((n, cb), dtype=torch.int64, device=dev)
        fn(P(tnb), P(mask), n, ctypes.c_float(thr)); ref.ref_sync()
        m = mask.cpu().numpy().view(np.uint64)
        remv = np.zeros(cb, np.uint64); keep = []
        for i in range(n):
            if not (int(remv[i // 64]) >> (i % 64)) & 1:
                keep.append(i); remv[i // 64:] |= m[i, i // 64:]
        keeps[name] = np.asarray(keep, np.int64)
    ref.ref_sync()
    np.savez_compressed(os.path.join(out_dir, "ref_kernels_iou3d.npz"), a=a, b=b, overlap=ov.cpu().numpy(), iou=iou.cpu().numpy(),
                        nms_boxes=nb, keep_rot=keeps["rot"], thr_rot=np.float32(0.25), keep_normal=keeps["normal"], thr_normal=np.float32(0.5))

    # --- roiaware: points in boxes + pool
    B, T, M = 2, 40, 4000
    boxes = np.stack([rand_boxes(rng, T, 12) for _ in range(B)])
    pts = rng.uniform([-14, -14, -2.5], [14, 14, 1.5], (B, M, 3)).astype(np.float32)
    out = torch.full((B, M), -1, dtype=torch.int32, device=dev)
    tb_, tp_ = cu(boxes), cu(pts)
    ref.ref_points_in_boxes(B, T, M, P(tb_), P(tp_), P(out)); ref.ref_sync()
    rois, feat = boxes[0][:12], rng.normal(size=(M, 8)).astype(np.float32)
    res = {}
    for method, m in (("max", 0), ("avg", 1)):
        pooled = torch.zeros((12, 4, 4, 4, 8), device=dev)
        argmax = torch.zeros((12, 4, 4, 4, 8), dtype=torch.int32, device=dev)
        pidx = torch.zeros((12, 4, 4, 4, 10), dtype=torch.int32, device=dev)
        tr, tp0, tf = cu(rois), cu(pts[0]), cu(feat)
        ref.ref_roiaware_pool3d(12, M, 8, 10, 4, 4, 4, P(tr), P(tp0), P(tf), P(argmax), P(pidx), P(pooled), m); ref.ref_sync()
        res[method] = (pooled.cpu().numpy(), argmax.cpu().numpy(), pidx.cpu().numpy())
    np.savez_compressed(os.path.join(out_dir, "ref_kernels_roiaware.npz"), boxes=boxes, pts=pts, pib=out.cpu().numpy(), rois=rois, feat=feat,
                        pooled_max=res["max"][0], argmax=res["max"][1], pidx=res["max"][2], pooled_avg=res["avg"][0])

    # --- pointnet2
    xyz_cnt, new_cnt = np.array([1500, 900], np.int32), np.array([200, 150], np.int32)
    xyz = rng.uniform(-6, 6, (2400, 3)).astype(np.float32)
    new_xyz = np.concatenate([xyz[:200] + np.float32(0.03), xyz[1500:1650] + np.float32(0.02)])
    idx = torch.zeros((350, 16), dtype=torch.int32, device=dev)
    t1, t2, t3, t4 = cu(new_xyz), cu(new_cnt), cu(xyz), cu(xyz_cnt)
    ref.ref_ball_query(2, 350, ctypes.c_float(0.8), 16, P(t1), P(t2), P(t3), P(t4), P(idx)); ref.ref_sync()
    fpts = rng.uniform(-10, 10, (2, 3000, 3)).astype(np.float32)
    temp = torch.full((2, 3000), 1e10, device=dev)
    fidx = torch.zeros((2, 256), dtype=torch.int32, device=dev)
    tf_ = cu(fpts)
    ref.ref_fps(2, 3000, 256, P(tf_), P(temp), P(fidx)); ref.ref_sync()
    fpts2 = rng.uniform(-10, 10, (1, 700, 3)).astype(np.float32)   # n < 1024: reference block = 512
    temp2 = torch.full((1, 700), 1e10, device=dev)
    fidx2 = torch.zeros((1, 64), dtype=torch.int32, device=dev)
    tf2 = cu(fpts2)
    ref.ref_fps(1, 700, 64, P(tf2), P(temp2), P(fidx2)); ref.ref_sync()
    uc, kc = np.array([500, 300], np.int32), np.array([120, 90], np.int32)
    unknown, known = rng.uniform(-6, 6, (800, 3)).astype(np.float32), rng.uniform(-6, 6, (210, 3)).astype(np.float32)
    d2 = torch.zeros((800, 3), device=dev); nidx = torch.zeros((800, 3), dtype=torch.int32, device=dev)
    a1, a2, a3, a4 = cu(unknown), cu(uc), cu(known), cu(kc)
    ref.ref_three_nn(2, 800, 210, P(a1), P(a2), P(a3), P(a4), P(d2), P(nidx)); ref.ref_sync()
    np.savez_compressed(os.path.join(out_dir, "ref_kernels_pointnet2.npz"), xyz=xyz, xyz_cnt=xyz_cnt, new_xyz=new_xyz, new_cnt=new_cnt,
                        bq_idx=idx.cpu().numpy(), bq_radius=np.float32(0.8), fps_pts=fpts, fps_idx=fidx.cpu().numpy(),
                        fps_pts2=fpts2, fps_idx2=fidx2.cpu().numpy(), unknown=unknown, uc=uc, known=known, kc=kc,
                        nn_d2=d2.cpu().numpy(), nn_idx=nidx.cpu().numpy())
    print("golden written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
