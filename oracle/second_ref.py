"""CPU restatement of the SECOND forward + CRB stage-1 record (TEST INFRASTRUCTURE / CPU baseline ONLY).

Follows pcdet/models/detectors/second_net.py:9-22 through the reference modules: MeanVFE (mean_vfe.py:14-31),
VoxelBackBone8x (spconv_backbone.py:69-180; spconv 'Native' CPU algorithm = gather -> torch.mm -> index_add),
HeightCompression (height_compression.py:10-26), BaseBEVBackbone (base_bev_backbone.py:81-112), AnchorHeadSingle +
generate_predicted_boxes (anchor_head_single.py:41-76, anchor_head_template.py:238-285, box_coder_utils.py:45-77),
post_processing / class_agnostic_nms (detector3d_template.py:186-409, model_nms_utils.py:6-25, iou3d_nms_utils.py:84-99),
per-box density (detector3d_template.py:379-387) and the stage-1 entropy (crb_sampling.py:86-100).
Weights come from a state_dict with the reference's parameter names.
"""
import numpy as np
import torch
import torch.nn.functional as F

from . import boxes as ob, crb as oc, spconv_ref, voxel

# (prefix, conv type, indice_key, ksize, stride, padding)
BACKBONE_LAYERS = [
    ("conv_input", "subm", "subm1", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv1.0", "subm", "subm1", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv2.0", "spconv", "spconv2", (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("conv2.1", "subm", "subm2", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv2.2", "subm", "subm2", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv3.0", "spconv", "spconv3", (3, 3, 3), (2, 2, 2), (1, 1, 1)),
    ("conv3.1", "subm", "subm3", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv3.2", "subm", "subm3", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv4.0", "spconv", "spconv4", (3, 3, 3), (2, 2, 2), (0, 1, 1)),
    ("conv4.1", "subm", "subm4", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv4.2", "subm", "subm4", (3, 3, 3), (1, 1, 1), (1, 1, 1)),
    ("conv_out", "spconv", "spconv_down2", (3, 1, 1), (2, 1, 1), (0, 0, 0)),
]


# when True, every BatchNorm first overwrites its running statistics in `sd` with the statistics of its input (the CPU twin
# of crb3d.second.calibrate_batchnorm: one train-mode forward with momentum 1), then normalises with them
CALIBRATE_BN = False


def _bn_eval(x, sd, prefix, eps=1e-3):
    if CALIBRATE_BN:
        sd[prefix + ".running_mean"] = x.mean(0)
        sd[prefix + ".running_var"] = x.var(0, unbiased=True)
    w, b, rm, rv = (sd[prefix + s].float() for s in (".weight", ".bias", ".running_mean", ".running_var"))
    return (x - rm) / torch.sqrt(rv + eps) * w + b


def _bn2d(x, sd, prefix):
    if CALIBRATE_BN:
        sd[prefix + ".running_mean"] = x.mean((0, 2, 3))
        sd[prefix + ".running_var"] = x.var((0, 2, 3), unbiased=True)
    return F.batch_norm(x, sd[prefix + ".running_mean"], sd[prefix + ".running_var"], sd[prefix + ".weight"], sd[prefix + ".bias"],
                        False, 0.0, 1e-3)


def backbone3d(sd, feats, coords, batch_size, sparse_shape, collect=None):
    x = torch.as_tensor(feats).float()
    coords = np.asarray(coords)
    shape = list(sparse_shape)
    books = {}
    for prefix, kind, key, ks, st, pad in BACKBONE_LAYERS:
        w = sd["backbone_3d.%s.0.weight" % prefix].float()
        if key not in books:
            if kind == "subm":
                books[key] = (coords, shape, spconv_ref.subm_rulebook(coords, shape, ks))
            else:
                oc_, osh, nbr, _ = spconv_ref.sparse_rulebook(coords, batch_size, shape, ks, st, pad)
                books[key] = (oc_, osh, nbr)
        coords, shape, nbr = books[key]
        x = spconv_ref.conv_forward(x, nbr, w)
        x = torch.relu(_bn_eval(x, sd, "backbone_3d.%s.1" % prefix))
        if collect is not None:
            collect[prefix] = (x.clone(), coords.copy(), list(shape))
    return x, coords, shape


def bev_head(sd, spatial, cfg):
    x = spatial
    ups = []
    for i, (n_layers, stride) in enumerate(zip(cfg["layer_nums"], cfg["layer_strides"])):
        p = "backbone_2d.blocks.%d" % i
        x = F.conv2d(F.pad(x, (1, 1, 1, 1)), sd[p + ".1.weight"].float(), stride=stride)
        x = torch.relu(_bn2d(x, sd, p + ".2"))
        for k in range(n_layers):
            j = 4 + 3 * k
            x = F.conv2d(x, sd["%s.%d.weight" % (p, j)].float(), padding=1)
            x = torch.relu(_bn2d(x, sd, "%s.%d" % (p, j + 1)))
        d = "backbone_2d.deblocks.%d" % i
        u = F.conv_transpose2d(x, sd[d + ".0.weight"].float(), stride=cfg["upsample_strides"][i])
        ups.append(torch.relu(_bn2d(u, sd, d + ".1")))
    x = torch.cat(ups, 1)
    B = x.shape[0]
    cls = F.conv2d(x, sd["dense_head.conv_cls.weight"], sd["dense_head.conv_cls.bias"]).permute(0, 2, 3, 1).reshape(B, -1, len(cfg["class_names"]))
    box = F.conv2d(x, sd["dense_head.conv_box.weight"], sd["dense_head.conv_box.bias"]).permute(0, 2, 3, 1).reshape(B, -1, 7)
    dr = F.conv2d(x, sd["dense_head.conv_dir_cls.weight"], sd["dense_head.conv_dir_cls.bias"]).permute(0, 2, 3, 1).reshape(B, -1, cfg["num_dir_bins"])
    return cls, box, dr


def decode_boxes(box, dr, anchors, cfg):
    """ResidualCoder.decode_torch + direction fix (anchor_head_template.py:262-278), all anchors."""
    a = anchors.unsqueeze(0)
    diag = torch.sqrt(a[..., 3] ** 2 + a[..., 4] ** 2)
    out = torch.empty_like(box)
    out[..., 0] = box[..., 0] * diag + a[..., 0]
    out[..., 1] = box[..., 1] * diag + a[..., 1]
    out[..., 2] = box[..., 2] * a[..., 5] + a[..., 2]
    out[..., 3:6] = torch.exp(box[..., 3:6]) * a[..., 3:6]
    rg = box[..., 6] + a[..., 6]
    period = 2 * np.pi / cfg["num_dir_bins"]
    val = rg - cfg["dir_offset"]
    dir_rot = val - torch.floor(val / period + cfg["dir_limit_offset"]) * period
    out[..., 6] = dir_rot + cfg["dir_offset"] + period * torch.max(dr, dim=-1)[1].to(box.dtype)
    return out


def score_frames(sd, cfg, frames, anchors, collect=None, threads=None, timers=None):
    """Full CPU path for a list of per-frame point arrays. Returns a list of per-frame records.
    timers (optional dict): accumulates wall seconds per stage (voxelize, sparse_backbone, dense, bev_head, post, density)."""
    import time as _time

    def tick(name, t0):
        if timers is not None:
            timers[name] = timers.get(name, 0.0) + (_time.perf_counter() - t0)
        return _time.perf_counter()
    if threads:
        torch.set_num_threads(threads)
    d = cfg["data"]
    if not CALIBRATE_BN:      # (calibration writes the new statistics into the caller's dict)
        sd = {k: v.detach().cpu() for k, v in sd.items()}
    B = len(frames)
    with torch.no_grad():
        t0 = _time.perf_counter()
        feats, coords, _, _ = voxel.voxelize_batch(frames, d["pc_range"], d["voxel_size"], d["max_pts"], d["max_voxels_test"])
        t0 = tick("voxelize", t0)
        x, c4, shape = backbone3d(sd, feats, coords, B, d["sparse_shape"], collect)
        t0 = tick("sparse_backbone", t0)
        dense = spconv_ref.dense(x, c4, B, shape)
        spatial = dense.view(B, -1, shape[1], shape[2])
        t0 = tick("dense", t0)
        if collect is not None:
            collect["spatial_features"] = spatial.clone()
        cls, box, dr = bev_head(sd, spatial, cfg)
        t0 = tick("bev_head", t0)
        if collect is not None:
            collect["cls_preds"], collect["box_preds"], collect["dir_cls_preds"] = cls, box, dr
        boxes_all = decode_boxes(box, dr, anchors, cfg)
        recs = []
        for b in range(B):
            sc = torch.sigmoid(cls[b])
            conf, lab = torch.max(sc, dim=-1)
            lab = lab + 1
            mask = conf >= cfg["score_thresh"]
            conf_m, box_m, lab_m = conf[mask], boxes_all[b][mask], lab[mask]
            if conf_m.numel():
                top, idx = torch.topk(conf_m, k=min(cfg["nms_pre_maxsize"], conf_m.shape[0]))
                cand = box_m[idx]
                keep = ob.nms_sorted(cand.numpy(), cfg["nms_thresh"])[: cfg["nms_post_maxsize"]]
                fb, fs, fl = cand[keep], top[keep], lab_m[idx][keep]
            else:
                fb, fs, fl = box_m[:0], conf_m[:0], lab_m[:0]
            t0 = tick("post", t0)
            dens, cnt, _ = ob.box_density(fb.numpy(), frames[b][:, :3]) if len(fb) else (np.zeros(0, np.float32), np.zeros(0, np.int32), None)
            recs.append(dict(boxes=fb.numpy(), scores=fs.numpy(), labels=fl.numpy(), density=dens, point_counts=cnt,
                             entropy=oc.label_entropy(fl.numpy(), len(cfg["class_names"]))))
            t0 = tick("density", t0)
    return recs
