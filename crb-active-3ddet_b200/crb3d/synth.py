"""Deterministic LiDAR-like synthetic frames (no dataset access in this environment).

KITTI-shaped ("K"): range/voxel/limits of tools/cfgs/dataset_configs/kitti_dataset.yaml:4,65-70 of the reference
(POINT_CLOUD_RANGE [0,-40,-3,70.4,40,1], VOXEL_SIZE [.05,.05,.1], 5 pts/voxel, 16000/40000 voxels).
Waymo-shaped ("W"): waymo_dataset.yaml:5,73-78 ([-75.2,-75.2,-2,75.2,75.2,4], [.1,.1,.15], 5 feats, 150000 voxels).
Generator recipe follows SURVEY.md 8(d): frame i uses np.random.default_rng(1000+i); 64 beams; rays hit a ground
plane or one of 10-40 random class-sized boxes; range noise N(0,0.02); intensity U(0,1).
"""
import numpy as np

KITTI = dict(
    name="kitti", pc_range=[0.0, -40.0, -3.0, 70.4, 40.0, 1.0], voxel_size=[0.05, 0.05, 0.1], n_feat=4, max_pts=5,
    max_voxels_train=16000, max_voxels_test=40000, sparse_shape=[41, 1600, 1408], target_points=20000,
    sensor_z=0.0, ground_z=-1.73, az_range=(-45.0, 45.0), az_step=0.25, n_beams=64, elev=(-24.8, 2.0),
)
WAYMO = dict(
    name="waymo", pc_range=[-75.2, -75.2, -2.0, 75.2, 75.2, 4.0], voxel_size=[0.1, 0.1, 0.15], n_feat=5, max_pts=5,
    max_voxels_train=150000, max_voxels_test=150000, sparse_shape=[41, 1504, 1504], target_points=160000,
    sensor_z=2.0, ground_z=0.0, az_range=(-180.0, 180.0), az_step=0.1, n_beams=64, elev=(-17.6, 2.4),
)
# anchor sizes (dx, dy, dz) of tools/cfgs/kitti_models/second.yaml:40-68
CLASS_SIZES = np.array([[3.9, 1.6, 1.56], [0.8, 0.6, 1.73], [1.76, 0.6, 1.73]], dtype=np.float64)
CLASS_MIX = np.array([0.6, 0.3, 0.1])


def _ray_boxes(origin, dirs, boxes):
    """Smallest positive hit range of each ray against yaw-rotated boxes (slab test in the box frame)."""
    n = len(dirs)
    best = np.full(n, np.inf)
    for b in boxes:
        c, s = np.cos(-b[6]), np.sin(-b[6])
        o = origin - b[:3]
        ox, oy = o[0] * c - o[1] * s, o[0] * s + o[1] * c
        dx, dy = dirs[:, 0] * c - dirs[:, 1] * s, dirs[:, 0] * s + dirs[:, 1] * c
        oo = np.array([ox, oy, o[2]])
        dd = np.stack([dx, dy, dirs[:, 2]], 1)
        half = b[3:6] / 2
        with np.errstate(divide="ignore", invalid="ignore"):
            t1 = (-half - oo) / dd
            t2 = (half - oo) / dd
        tmin = np.nanmax(np.minimum(t1, t2), axis=1)
        tmax = np.nanmin(np.maximum(t1, t2), axis=1)
        hit = (tmax >= tmin) & (tmin > 0.5)
        best = np.where(hit & (tmin < best), tmin, best)
    return best


def make_boxes(rng, cfg, n_boxes=None):
    """Random ground-standing boxes (n,8): x,y,z,dx,dy,dz,yaw,class(1..3)."""
    n = int(rng.integers(10, 41)) if n_boxes is None else n_boxes
    cls = rng.choice(3, size=n, p=CLASS_MIX)
    size = CLASS_SIZES[cls] * rng.uniform(0.9, 1.1, size=(n, 3))
    r = cfg["pc_range"]
    x = rng.uniform(max(r[0], -60) + 4, min(r[3], 60) - 4, n)
    y = rng.uniform(max(r[1], -35) + 2, min(r[4], 35) - 2, n)
    z = cfg["ground_z"] + size[:, 2] / 2
    yaw = rng.uniform(-np.pi, np.pi, n)
    return np.concatenate([np.stack([x, y, z], 1), size, yaw[:, None], (cls + 1)[:, None]], 1)


def make_frame(index, cfg=KITTI, return_boxes=False):
    """One synthetic frame: (N, n_feat) float32 points inside POINT_CLOUD_RANGE (x,y,z,intensity[,elongation])."""
    rng = np.random.default_rng(1000 + index)
    boxes = make_boxes(rng, cfg)
    target = cfg["target_points"]
    n_beams = cfg["n_beams"]
    az = np.deg2rad(np.arange(cfg["az_range"][0], cfg["az_range"][1], cfg["az_step"]))
    # raise the beam count until the ray budget comfortably exceeds the target point count
    while n_beams * len(az) < target * 1.6:
        n_beams *= 2
    el = np.deg2rad(np.linspace(cfg["elev"][0], cfg["elev"][1], n_beams))
    A, E = np.meshgrid(az, el)
    dirs = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], -1).reshape(-1, 3)
    origin = np.array([0.0, 0.0, cfg["sensor_z"]])
    with np.errstate(divide="ignore"):
        t_ground = np.where(dirs[:, 2] < -1e-6, (cfg["ground_z"] - origin[2]) / dirs[:, 2], np.inf)
    t = np.minimum(t_ground, _ray_boxes(origin, dirs, boxes))
    ok = np.isfinite(t) & (t < 120.0)
    t = t[ok] + rng.normal(0.0, 0.02, ok.sum())
    p = origin + dirs[ok] * t[:, None]
    r = cfg["pc_range"]
    inside = (p[:, 0] >= r[0]) & (p[:, 0] < r[3]) & (p[:, 1] >= r[1]) & (p[:, 1] < r[4]) & (p[:, 2] >= r[2]) & (p[:, 2] < r[5])
    p = p[inside]
    want = int(target + rng.integers(-target // 10, target // 10 + 1))
    if len(p) > want:
        p = p[np.sort(rng.choice(len(p), want, replace=False))]
    feats = [p, rng.uniform(0, 1, (len(p), 1))]
    if cfg["n_feat"] == 5:
        feats.append(rng.uniform(0, 1, (len(p), 1)))
    pts = np.concatenate(feats, 1).astype(np.float32)
    return (pts, boxes.astype(np.float32)) if return_boxes else pts


def make_batch(indices, cfg=KITTI):
    """Stacked batch like DatasetTemplate.collate_batch (pcdet/datasets/dataset.py:180-186): points (N, 1+C) with the
    batch index in column 0, plus int32 frame offsets (B+1)."""
    frames = [make_frame(i, cfg) for i in indices]
    offs = np.zeros(len(frames) + 1, dtype=np.int32)
    offs[1:] = np.cumsum([len(f) for f in frames])
    pts = np.concatenate([np.concatenate([np.full((len(f), 1), b, np.float32), f], 1) for b, f in enumerate(frames)])
    return pts, offs, frames
