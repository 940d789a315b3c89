"""Strategy base class of the active-learning loop: same constructor, attributes and methods as
pcdet/query_strategies/strategy.py:5-81, which `select_active_labels` (pcdet/utils/active_training_utils.py:252-293) relies on:
`pairs`, `labelled_set`, `unlabelled_set`, `save_points`, `save_active_labels`, `update_dashboard`, `query`.

Differences that do not change results: wandb is optional (the reference imports it unconditionally; without it the dashboard
values are kept in `self.dashboard_log` instead), and loaders without a `.dataset` (the dict-of-frames form the tests and the
benchmark use) are accepted - `pairs` is then built from the frame ids.
"""
import os
import pickle

try:  # the reference logs to wandb (strategy.py:3); not a dependency of the hot path
    import wandb  # noqa: F401
    _HAVE_WANDB = getattr(wandb, "run", None) is not None or hasattr(wandb, "log")
except Exception:  # pragma: no cover
    wandb = None
    _HAVE_WANDB = False


def _cfg_get(cfg, path, default=None):
    cur = cfg
    for key in path.split("."):
        if cur is None:
            return default
        cur = cur.get(key, None) if isinstance(cur, dict) else getattr(cur, key, None)
    return default if cur is None else cur


class Strategy(object):
    def __init__(self, model, labelled_loader, unlabelled_loader, rank, active_label_dir, cfg):
        self.cfg = cfg
        self.active_label_dir = active_label_dir
        self.rank = rank
        self.model = model
        self.labelled_loader = labelled_loader
        self.unlabelled_loader = unlabelled_loader
        self.labelled_set = getattr(labelled_loader, "dataset", None)
        self.unlabelled_set = getattr(unlabelled_loader, "dataset", None)
        self.bbox_records = {}
        self.point_measures = ["mean", "median", "variance"]
        for met in self.point_measures:
            setattr(self, "{}_point_records".format(met), {})
        self.dashboard_log = []
        ds = self.unlabelled_set
        if ds is not None and _cfg_get(cfg, "DATA_CONFIG.DATASET") == "KittiDataset" and hasattr(ds, "sample_id_list"):
            self.pairs = list(zip(ds.sample_id_list, ds.kitti_infos))            # strategy.py:22-23
        elif ds is not None and hasattr(ds, "frame_ids"):
            self.pairs = list(zip(ds.frame_ids, ds.infos))                        # strategy.py:24-25
        elif isinstance(unlabelled_loader, dict):
            self.pairs = [(fid, {"frame_id": fid}) for fid in unlabelled_loader]
        else:
            self.pairs = []

    def save_points(self, frame_id, batch_dict):
        """strategy.py:27-38: per-frame box / point statistics shown on the dashboard."""
        self.bbox_records[frame_id] = batch_dict["num_bbox"]
        self.mean_point_records[frame_id] = batch_dict["mean_points"]
        self.median_point_records[frame_id] = batch_dict["median_points"]
        self.variance_point_records[frame_id] = batch_dict["variance_points"]

    def _log(self, payload, step=None):
        self.dashboard_log.append((step, payload))
        if _HAVE_WANDB and getattr(wandb, "run", None) is not None:
            wandb.log(payload, step=step)

    def update_dashboard(self, cur_epoch=None, accumulated_iter=None):
        """strategy.py:42-63."""
        if not getattr(self, "selected_bbox", None):
            return
        classes = list(self.selected_bbox[0].keys())
        total_bbox = 0
        for cls_idx in classes:
            num_cls_bbox = sum([i[cls_idx] for i in self.selected_bbox])
            self._log({"active_selection/num_bbox_{}".format(cls_idx): num_cls_bbox}, step=accumulated_iter)
            total_bbox += num_cls_bbox
            for met in self.point_measures:
                sel = getattr(self, "selected_{}_points".format(met))
                stats_point = (sum([i[cls_idx] for i in sel]) / len(sel)) if num_cls_bbox else 0
                self._log({"active_selection/{}_points_{}".format(met, cls_idx): stats_point}, step=accumulated_iter)
        self._log({"active_selection/total_bbox_selected": total_bbox}, step=accumulated_iter)

    def save_active_labels(self, selected_frames=None, grad_embeddings=None, cur_epoch=None):
        """strategy.py:66-81: pickles the selection (and optionally the gradient embeddings) under active_label_dir."""
        if selected_frames is not None:
            self.selected_bbox = [self.bbox_records[i] for i in selected_frames]
            for met in self.point_measures:
                setattr(self, "selected_{}_points".format(met),
                        [getattr(self, "{}_point_records".format(met))[i] for i in selected_frames])
            if self.active_label_dir is not None:
                path = os.path.join(self.active_label_dir, "selected_frames_epoch_{}_rank_{}.pkl".format(cur_epoch, self.rank))
                with open(path, "wb") as f:
                    pickle.dump({"frame_id": selected_frames, "selected_mean_points": self.selected_mean_points,
                                 "selected_bbox": self.selected_bbox, "selected_median_points": self.selected_median_points,
                                 "selected_variance_points": self.selected_variance_points}, f)
                print("successfully saved selected frames for epoch {} for rank {}".format(cur_epoch, self.rank))
        if grad_embeddings is not None and self.active_label_dir is not None:
            with open(os.path.join(self.active_label_dir, "grad_embeddings_epoch_{}.pkl".format(cur_epoch)), "wb") as f:
                pickle.dump(grad_embeddings, f)
            print("successfully saved grad embeddings for epoch {}".format(cur_epoch))

    def query(self, leave_pbar=True, cur_epoch=None):
        pass
