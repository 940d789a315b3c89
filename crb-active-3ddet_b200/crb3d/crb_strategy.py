"""CRB acquisition strategy on the crb3d kernels: same constructor and `query(leave_pbar, cur_epoch)` contract as
pcdet/query_strategies/crb_sampling.py:20-342 (registered as 'crb' in pcdet/query_strategies/__init__.py:13-29).

Stage 1 (concise label sampling, :72-121)  - forward + post-processing + per-frame label entropy on the device
          (PoolScorer.score_pool: frames sharded over ranks, one all-gather of the records); ranking by the reference's
          stable-sort-then-reverse rule.
Stage 2 (representative prototypes, :134-226) - gradient embedding per shortlisted frame, pairwise squared distances on
          the device (crb3d_pairwise_sqdist_f64), k-means++ seeding restated with sklearn's RandomState(0) stream.
          * detectors with a RoI head: the reference's embedding, roi_head.shared_fc_layer[4].weight.grad (:194-207);
          * SECOND (no RoI head - the reference's CRB cannot run it, SURVEY.md 8a note G): the in-repo precedent of
            BADGE (badge_sampling.py:88-91,157-168), the gradient of dense_head.conv_cls.weight under the sigmoid focal
            loss with the arg-max hypothetical labels. conv_cls is the LAST layer, so its weight gradient is the closed
            form  dW = delta^T X  (delta = dLoss/dlogits, X = the 512-channel BEV map): no backward pass is needed.
Stage 3 (greedy density balancing, :247-338) - uniform prior from the WHOLE pool, greedy KDE/KL selection on the device
          (crb3d_kde_greedy).
Quirks that define results are kept: pseudo-count 1 for absent classes, `BANDWDITH` typo (bandwidth is always 5 unless
that misspelt key is set), int() truncation of the density quantiles, strict-'>' first maximum.
"""
import types

import numpy as np
import torch

from . import crb_host, ops
from .strategy import Strategy, _cfg_get as _get


def sigmoid_focal_loss_grad(logits, labels, alpha=0.25, gamma=2.0):
    """d/dlogits of pcdet's SigmoidFocalClassificationLoss (loss_utils.py:9-72) summed over anchors with the anchor-head
    weighting of anchor_head_template.py:95-128 (positives + negatives weighted 1, normalised by the positive count).
    labels: (A,) int, 0 = background, c > 0 = class c (one-hot column c-1). Returns delta (A, n_class)."""
    n_class = logits.shape[-1]
    target = torch.zeros_like(logits)
    pos = labels > 0
    target[pos, (labels[pos] - 1).long()] = 1.0
    p = torch.sigmoid(logits)
    alpha_w = target * alpha + (1 - target) * (1 - alpha)
    pt = target * (1.0 - p) + (1.0 - target) * p
    focal = alpha_w * torch.pow(pt, gamma)
    # bce = max(x,0) - x*t + log1p(exp(-|x|)); loss = focal * bce ; closed-form derivative (no autograd graph kept)
    bce = torch.clamp(logits, min=0) - logits * target + torch.log1p(torch.exp(-torch.abs(logits)))
    dbce = p - target
    dpt = (1.0 - 2.0 * target) * p * (1.0 - p)
    dfocal = alpha_w * gamma * torch.pow(pt, gamma - 1.0) * dpt
    norm = torch.clamp(pos.sum().float(), min=1.0)
    return (dfocal * bce + focal * dbce) / norm


class CRBSampling(Strategy):
    """A `Strategy` (crb3d/strategy.py = pcdet/query_strategies/strategy.py:5-81): `build_strategy('crb', ...)` of the reference
    returns this class when crb3d.dropin is installed (alias pcdet.query_strategies.crb_sampling)."""

    def __init__(self, model, labelled_loader, unlabelled_loader, rank, active_label_dir, cfg):
        super(CRBSampling, self).__init__(model, labelled_loader, unlabelled_loader, rank, active_label_dir, cfg)
        self.k1 = _get(cfg, "ACTIVE_TRAIN.ACTIVE_CONFIG.K1", 5)
        self.k2 = _get(cfg, "ACTIVE_TRAIN.ACTIVE_CONFIG.K2", 3)
        self.bandwidth = _get(cfg, "ACTIVE_TRAIN.ACTIVE_CONFIG.BANDWDITH", 5)   # sic: the reference reads the typo
        self.prototype = _get(cfg, "ACTIVE_TRAIN.ACTIVE_CONFIG.CLUSTERING", "kmeans++")
        self.select_nums = int(_get(cfg, "ACTIVE_TRAIN.SELECT_NUMS", 100))
        self.alpha = 0.95
        self.last_stage = {}
        self._gt_boxes = {}

    # ------------------------------------------------------------------------------------------------ stage 1
    def collect_pool(self):
        """{frame_id: points (n, C) float32}. The loader yields pcdet-style batches: 'frame_id' (B,), 'points'
        (N, 1+C) with the batch index in column 0 (pcdet/datasets/dataset.py:180-186), or a dict of frames directly."""
        if isinstance(self.unlabelled_loader, dict):
            return dict(self.unlabelled_loader)
        frames = {}
        for batch in self.unlabelled_loader:
            pts = np.asarray(batch["points"])
            for b, fid in enumerate(batch["frame_id"]):
                frames[fid] = pts[pts[:, 0] == b][:, 1:].astype(np.float32)
                if "gt_boxes" in batch:      # dashboard statistics of strategy.save_points (detector3d_template.py:240-267)
                    self._gt_boxes[fid] = np.asarray(batch["gt_boxes"][b], dtype=np.float32)
        return frames

    def _point_stats(self, fid, points, class_names):
        """num_bbox / mean / median / variance of the points inside the frame's ground-truth boxes per class
        (detector3d_template.py:240-267; zeros when the loader carries no gt_boxes, e.g. a raw unlabelled pool)."""
        stats = {k: {c: 0 for c in class_names} for k in ("num_bbox", "mean_points", "median_points", "variance_points")}
        gt = self._gt_boxes.get(fid)
        if gt is None or len(gt) == 0:
            return stats
        dev = next(self.model.parameters()).device
        pts = torch.from_numpy(np.ascontiguousarray(points[:, :3], dtype=np.float32)).to(dev).view(1, -1, 3)
        for ci, cname in enumerate(class_names):
            sel = gt[gt[:, -1] == ci + 1]
            stats["num_bbox"][cname] = int(len(sel))
            if len(sel) == 0:
                continue
            idx = ops.points_in_boxes(torch.from_numpy(np.ascontiguousarray(sel[:, :7])).to(dev).view(1, -1, 7), pts).view(-1)
            cnt = torch.bincount(idx[idx >= 0].long(), minlength=1).float()
            cnt = cnt[cnt > 0]          # the reference counts the boxes that appear in torch.unique (boxes that own a point)
            if cnt.numel():
                stats["mean_points"][cname] = float(cnt.mean())
                stats["median_points"][cname] = float(cnt.median())
                stats["variance_points"][cname] = float(cnt.var(unbiased=False))
        return stats

    def stage1(self, scorer, frames):
        ids = list(frames.keys())
        recs = scorer.score_pool([frames[i] for i in ids], ids)
        entropies = [recs[i]["entropy"] for i in ids]
        shortlist = crb_host.shortlist_by_entropy(ids, entropies, int(self.k1 * self.select_nums))
        return recs, shortlist

    # ------------------------------------------------------------------------------------------------ stage 2
    @torch.no_grad()
    def second_gradient_embedding(self, scorer, points):
        """(n_loc*n_class*512,) embedding of one frame for a SECOND-style detector (see module docstring)."""
        dev = scorer.device
        pts = torch.from_numpy(np.ascontiguousarray(points[:, -scorer.n_feat:], dtype=np.float32)).to(dev)
        offs = torch.tensor([0, pts.shape[0]], dtype=torch.int32, device=dev)
        bd = self.model.forward_features(pts, offs, 1)
        logits = bd["cls_preds"][0]                                        # (A, n_class), A = H*W*n_loc
        labels = torch.argmax(logits, dim=-1)                               # BADGE's hypothetical labels (0 = background)
        delta = sigmoid_focal_loss_grad(logits, labels)                     # (A, n_class)
        x = bd["spatial_features_2d"][0].permute(1, 2, 0)                   # (H, W, 512)
        n_loc, nc = self.model.dense_head.n_loc, logits.shape[-1]
        d = delta.view(x.shape[0] * x.shape[1], n_loc * nc)                 # channel = type*n_class + class (conv_cls rows)
        return (d.t() @ x.reshape(-1, x.shape[-1])).reshape(-1)             # conv_cls.weight.grad, flattened

    def stage2(self, scorer, frames, shortlist, embedding_fn=None):
        fn = embedding_fn or (lambda pts: self.second_gradient_embedding(scorer, pts))
        emb = torch.stack([fn(frames[fid]).float() for fid in shortlist], 0)
        n_clusters = int(self.select_nums * self.k2)
        if self.prototype != "kmeans++":
            raise NotImplementedError("only the paper's kmeans++ prototype selection is on the hot path")
        D = ops.pairwise_sqdist(emb)
        idx = crb_host.kmeans_plusplus_indices(D.cpu().numpy(), n_clusters, seed=0)
        self.last_stage["embeddings"] = emb
        return [shortlist[i] for i in idx]

    # ------------------------------------------------------------------------------------------------ stage 3
    def stage3(self, scorer, recs, prototypes, num_class):
        dev = scorer.device
        density_all = torch.from_numpy(np.concatenate([r["density"] for r in recs.values()]))
        label_all = torch.from_numpy(np.concatenate([r["labels"] for r in recs.values()]))
        axis, prior = crb_host.build_prior(density_all, label_all, num_class, self.alpha)
        dens = [recs[f]["density"] for f in prototypes]
        labs = [recs[f]["labels"] for f in prototypes]
        off = np.zeros(len(prototypes) + 1, np.int32)
        off[1:] = np.cumsum([len(d) for d in dens])
        order, scores = crb_host.greedy_density_balance(
            torch.from_numpy(np.concatenate(dens).astype(np.float32)).to(dev),
            torch.from_numpy(np.concatenate(labs).astype(np.int32)).to(dev), torch.from_numpy(off).to(dev), num_class,
            axis, prior, self.bandwidth, self.select_nums)
        self.last_stage["stage3_scores"] = scores
        return [prototypes[i] for i in order]

    # ------------------------------------------------------------------------------------------------ RoI-head detectors
    @staticmethod
    def _to_device(batch, dev):
        """pcdet.models.load_data_to_gpu (models/__init__.py:24-35) for the keys of a LiDAR batch."""
        out = {}
        for k, v in batch.items():
            if isinstance(v, np.ndarray) and k not in ("frame_id", "metadata", "calib", "sample_id_list"):
                out[k] = torch.from_numpy(v).float().to(dev)
            else:
                out[k] = v
        return out

    @staticmethod
    def enable_dropout(model):
        """crb_sampling.py:38-45: dropout layers stay stochastic at test time (Monte-Carlo rounds of the RoI head)."""
        n = 0
        for m in model.modules():
            if m.__class__.__name__.startswith("Dropout"):
                m.train()
                n += 1
        return n

    def query_roi_head(self, leave_pbar=True, cur_epoch=None, grad_batches=None):
        """The reference's query() for a detector with a RoI head (PV-RCNN), crb_sampling.py:48-342, step for step:
        stage 1  eval forward with MC dropout over the unlabelled loader -> per frame label entropy, hypothetical labels
                 (batch_rcnn_cls / batch_rcnn_reg), box labels and point densities; shortlist K1*N_r by entropy;
        stage 2  train-mode forward per shortlisted frame (batch size 1), RoI-head losses against the hypothetical labels,
                 gradient of shared_fc_layer[4].weight (crb3d.pvrcnn.roi_head_gradient_embedding), k-means++ (K2*N_r);
        stage 3  greedy KDE / KL density balancing on the device.
        `unlabelled_loader` yields pcdet batches (numpy or CUDA tensors); `grad_batches` (optional): {frame_id: batch of
        ONE frame} for stage 2 - the reference rebuilds a training dataloader there (build_active_dataloader, :150-160);
        without it the frames are cut out of the stage-1 batches."""
        from . import pvrcnn
        model = self.model
        dev = next(model.parameters()).device
        class_names = list(getattr(getattr(self.labelled_loader, "dataset", None), "class_names", [])) or \
            list(_get(self.cfg, "CLASS_NAMES", ["Car", "Pedestrian", "Cyclist"]))
        num_class = len(class_names)
        model.eval()
        self.enable_dropout(model)
        ents, cls_h, reg_h, dens, labs, singles = {}, {}, {}, {}, {}, {}
        for batch in self.unlabelled_loader:
            b = self._to_device(batch, dev)
            with torch.no_grad():
                pred_dicts, _ = model(dict(b))
            for i, p in enumerate(pred_dicts):
                fid = batch["frame_id"][i]
                fid = fid.item() if hasattr(fid, "item") else fid
                self.save_points(fid, p)
                lab = p["pred_labels"]
                ents[fid] = float(ops.label_entropy(lab.int().contiguous(), torch.tensor([0, lab.numel()], dtype=torch.int32, device=dev),
                                                    num_class)[0]) if lab.numel() else 0.0
                cls_h[fid], reg_h[fid] = p["batch_rcnn_cls"], p["batch_rcnn_reg"]
                dens[fid], labs[fid] = p["pred_box_unique_density"].float(), lab
                if grad_batches is None:
                    singles[fid] = self._single_frame(b, i)
        ids = list(ents.keys())
        self.last_stage["label_histogram"] = torch.bincount(torch.cat([labs[f].long().view(-1) for f in ids]), minlength=num_class + 1).tolist()
        shortlist = crb_host.shortlist_by_entropy(ids, [ents[i] for i in ids], int(self.k1 * self.select_nums))
        # ---- stage 2
        model.train()
        emb, index = [], []
        for fid in [p[0] for p in self.pairs if p[0] in set(shortlist)] or shortlist:      # the reference walks self.pairs (:141-146)
            fb = grad_batches[fid] if grad_batches is not None else singles[fid]
            emb.append(pvrcnn.roi_head_gradient_embedding(model, self._to_device(fb, dev), cls_h[fid], reg_h[fid]).float())
            index.append(fid)
        emb = torch.stack(emb, 0)
        if self.prototype != "kmeans++":
            raise NotImplementedError("only the paper's kmeans++ prototype selection is on the hot path")
        n_clusters = min(int(self.select_nums * self.k2), len(index))
        sel = crb_host.kmeans_plusplus_indices(ops.pairwise_sqdist(emb).cpu().numpy(), n_clusters, seed=0)
        prototypes = [index[i] for i in sel]
        # ---- stage 3
        recs = {f: dict(entropy=ents[f], labels=labs[f].cpu().numpy().astype(np.int64), density=dens[f].cpu().numpy()) for f in ids}
        scorer = types.SimpleNamespace(device=dev)
        selected = self.stage3(scorer, recs, prototypes, num_class)
        self.last_stage.update(dict(records=recs, shortlist=shortlist, prototypes=prototypes, embeddings=emb))
        model.eval()
        return selected

    @staticmethod
    def _single_frame(b, i):
        """Frame i of a device batch as a batch of one (points / voxels rows of that frame, batch index reset to 0)."""
        out = {"batch_size": 1, "frame_id": np.asarray([b["frame_id"][i]])}
        pm = b["points"][:, 0] == i
        out["points"] = torch.cat([torch.zeros_like(b["points"][pm][:, :1]), b["points"][pm][:, 1:]], 1).contiguous()
        if "voxel_coords" in b:
            vm = b["voxel_coords"][:, 0] == i
            vc = b["voxel_coords"][vm].clone()
            vc[:, 0] = 0
            out.update(voxels=b["voxels"][vm], voxel_num_points=b["voxel_num_points"][vm], voxel_coords=vc)
        if "gt_boxes" in b:
            out["gt_boxes"] = b["gt_boxes"][i:i + 1]
        return out

    # ------------------------------------------------------------------------------------------------ query
    def query(self, leave_pbar=True, cur_epoch=None, scorer=None, embedding_fn=None):
        if hasattr(self.model, "roi_head") and not isinstance(self.unlabelled_loader, dict):
            return self.query_roi_head(leave_pbar, cur_epoch)
        from .scorer import PoolScorer
        if scorer is None:
            dev = next(self.model.parameters()).device
            scorer = PoolScorer(self.model, dev, batch_size=4)
        class_names = list(self.model.cfg["class_names"]) if hasattr(self.model, "cfg") else list(self.labelled_loader.dataset.class_names)
        num_class = len(class_names)
        self.model.eval()
        if hasattr(self.model, "ensure_inference_current"):   # the model may have been trained since the plan / graphs were built
            self.model.ensure_inference_current()
        frames = self.collect_pool()
        for fid, pts in frames.items():
            self.save_points(fid, self._point_stats(fid, pts, class_names))
        recs, shortlist = self.stage1(scorer, frames)
        prototypes = self.stage2(scorer, frames, shortlist, embedding_fn)
        selected = self.stage3(scorer, recs, prototypes, num_class)
        self.last_stage.update(dict(records=recs, shortlist=shortlist, prototypes=prototypes))
        self.model.eval()
        return selected
