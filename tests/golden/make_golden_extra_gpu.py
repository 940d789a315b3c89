"""Generates tests/golden/ref_kernels_extra.npz from the REFERENCE's own voxel_query / roipoint_pool3d / pointnet2_batch CUDA
kernels (oracle/_ref/libpcdet_ref_kernels_batch.so, compiled from /root/reference by oracle/build.py). Run on a GPU box:
    python tests/golden/make_golden_extra_gpu.py gpurun_out/golden
then copy the .npz into tests/golden/. The CPU (-m "not gpu") tests pin oracle/csrc/oracle.c against it."""
import ctypes
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from util import P, rand_boxes, ref_batch_kernels  # noqa: E402


def main(out_dir):
    os.makedirs(out_dir, exist_ok=True)
    ref = ref_batch_kernels()
    assert ref is not None, "oracle/_ref not built"
    dev = torch.device("cuda:0")
    cu = lambda a: torch.as_tensor(np.ascontiguousarray(a)).to(dev)
    rng = np.random.default_rng(4048)

    # --- voxel query
    B, Z, Y, X, n = 2, 8, 24, 24, 1500
    voxel, origin = np.array([0.5, 0.5, 0.5]), np.array([-6.0, -6.0, -2.0])
    table = np.full((B, Z, Y, X), -1, np.int32)
    xyz = []
    for b in range(B):
        flat = rng.choice(Z * Y * X, n, replace=False)
        z, y, x = np.unravel_index(flat, (Z, Y, X))
        xyz.append(((np.stack([x, y, z], 1) + rng.uniform(0.05, 0.95, (n, 3))) * voxel + origin).astype(np.float32))
        table[b, z, y, x] = np.arange(n) + b * n
    xyz = np.concatenate(xyz)
    M = 600
    new_xyz = rng.uniform([-6.2, -6.2, -2.2], [6.2, 6.2, 2.2], (M, 3)).astype(np.float32)
    zyx = np.clip(np.floor((new_xyz[:, ::-1].astype(np.float64) - origin[::-1]) / voxel[::-1]).astype(np.int64), -1, [Z, Y, X])
    coords = np.concatenate([rng.integers(0, B, (M, 1)), zyx], 1).astype(np.int32)
    idx = torch.zeros((M, 16), dtype=torch.int32, device=dev)
    t = [cu(new_xyz), cu(xyz), cu(coords), cu(table)]
    ref.refb_voxel_query(M, Z, Y, X, 16, ctypes.c_float(0.9), 1, 2, 2, P(t[0]), P(t[1]), P(t[2]), P(t[3]), P(idx))
    assert ref.refb_sync() == 0

    # --- roipoint pool
    N, Mb, C, S = 3000, 24, 6, 64
    pts = rng.uniform([-12, -12, -2.5], [12, 12, 1.5], (1, N, 3)).astype(np.float32)
    feat = rng.normal(size=(1, N, C)).astype(np.float32)
    boxes = rand_boxes(rng, Mb, 10)[None]
    boxes[0, 0, :3] = 400.0
    boxes[0, 1, 3:6] = [25.0, 25.0, 8.0]
    pooled = torch.zeros((1, Mb, S, 3 + C), device=dev)
    empty = torch.zeros((1, Mb), dtype=torch.int32, device=dev)
    u = [cu(pts), cu(boxes), cu(feat)]
    ref.refb_roipoint_pool3d(1, N, Mb, C, S, P(u[0]), P(u[1]), P(u[2]), P(pooled), P(empty))
    assert ref.refb_sync() == 0

    # --- pointnet2_batch: ball query + 3-NN
    bx = rng.uniform(-5, 5, (2, 1200, 3)).astype(np.float32)
    bq = np.concatenate([bx[:, :100] + np.float32(0.02), rng.uniform(-20, 20, (2, 28, 3)).astype(np.float32)], 1)
    bidx = torch.zeros((2, 128, 16), dtype=torch.int32, device=dev)
    v = [cu(bq), cu(bx)]
    ref.refb_ball_query(2, 1200, 128, ctypes.c_float(0.7), 16, P(v[0]), P(v[1]), P(bidx))
    d2 = torch.zeros((2, 128, 3), device=dev)
    nidx = torch.zeros((2, 128, 3), dtype=torch.int32, device=dev)
    ref.refb_three_nn(2, 128, 1200, P(v[0]), P(v[1]), P(d2), P(nidx))
    assert ref.refb_sync() == 0
    np.savez_compressed(os.path.join(out_dir, "ref_kernels_extra.npz"), vq_xyz=xyz, vq_table=table, vq_new_xyz=new_xyz, vq_coords=coords,
                        vq_idx=idx.cpu().numpy(), vq_radius=np.float32(0.9), vq_range=np.array([1, 2, 2]),
                        rp_pts=pts[0], rp_feat=feat[0], rp_boxes=boxes[0], rp_pooled=pooled[0].cpu().numpy(), rp_empty=empty[0].cpu().numpy(),
                        bq_xyz=bx, bq_new=bq, bq_idx=bidx.cpu().numpy(), bq_radius=np.float32(0.7), nn_d2=d2.cpu().numpy(),
                        nn_idx=nidx.cpu().numpy())
    print("golden written to", out_dir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "gpurun_out", "golden"))
