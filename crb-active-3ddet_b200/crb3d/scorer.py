"""Pool scoring front end (the public call a CRB user makes for stage 1): host frames in, per-frame score records out.

One process per GPU. Frames are sharded round-robin over ranks like the reference's eval DistributedSampler
(pcdet/datasets/__init__.py:26-46: frame i -> rank i mod W); there is no data-path collective - the only exchange is ONE
all-gather of the fixed-stride per-frame records at the end, because stage 3's prior needs every pool
frame's (label, density) pairs (crb_sampling.py:252-258). The reference itself never synchronises ranks in query().
"""
import numpy as np
import torch
import torch.distributed as dist

RECORD_FIELDS = ("entropy", "num_boxes", "labels", "density")


class PoolScorer(object):
    def __init__(self, model, device, batch_size=4):
        self.model = model
        self.device = device
        self.batch_size = batch_size
        self.P = model.cfg["nms_post_maxsize"]
        self.n_feat = model.cfg["data"]["n_feat"]
        self._pinned = {}

    # -- host staging -------------------------------------------------------------------------------------------
    def stage_host(self, frames):
        """Packs a list of (n_i, C) float32 arrays into one pinned (N, C) tensor + int32 offsets (pinned)."""
        offs = np.zeros(len(frames) + 1, dtype=np.int32)
        offs[1:] = np.cumsum([len(f) for f in frames])
        pts = torch.empty((int(offs[-1]), self.n_feat), dtype=torch.float32).pin_memory()
        for i, f in enumerate(frames):
            pts[offs[i]:offs[i + 1]] = torch.from_numpy(np.ascontiguousarray(f[:, -self.n_feat:], dtype=np.float32))
        return pts, torch.from_numpy(offs).pin_memory(), int(max(len(f) for f in frames)) if frames else 0

    def to_device(self, staged):
        pts, offs, mx = staged
        return pts.to(self.device, non_blocking=True), offs.to(self.device, non_blocking=True), mx

    # -- scoring ------------------------------------------------------------------------------------------------
    def score_device(self, dev_batch):
        pts, offs, mx = dev_batch
        return self.model.score_batch(pts, offs, offs.numel() - 1, mx)

    def score_host(self, staged):
        """End-to-end call: pinned host points -> device -> kernels -> host record (numpy). With the whole-step CUDA
        graph (SECONDNet.enable_full_graph) this is one H2D copy into the static buffers, one graph launch and one D2H of
        the record; a batch that exceeds a static capacity is re-scored on the dynamic path."""
        fg = getattr(self.model, "_full_graph", None)
        if fg is not None and fg["B"] == staged[1].numel() - 1 and staged[0].shape[0] <= fg["cap"] and staged[2] <= fg["max_pts"]:
            rec = self.model.full_graph_replay(staged[0], staged[1])
            out = {k: rec[k].to("cpu", non_blocking=True) for k in RECORD_FIELDS + ("counts",)}
            torch.cuda.current_stream(self.device).synchronize()
            if bool((out["counts"].numpy() <= np.asarray(fg["caps"])).all()):
                return {k: out[k].numpy() for k in RECORD_FIELDS}
            dev = self.to_device(staged)   # capacity exceeded (never seen on KITTI-shaped clouds): dynamic path
            geom = self.model.geometry(dev[0], dev[1], dev[1].numel() - 1)
            rec = self.model.score_batch(dev[0], dev[1], dev[1].numel() - 1, dev[2], geom=geom)
        else:
            rec = self.score_device(self.to_device(staged))
        out = {k: rec[k].to("cpu", non_blocking=True) for k in RECORD_FIELDS}
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}

    def score_host_stream(self, staged_batches):
        """Throughput form of score_host for a sequence of staged (pinned) batches: consecutive batches alternate over
        the captured copies of the whole-step graph, each on its own stream - H2D of the points into that copy's static
        buffers, one graph launch, D2H of the record into pinned memory - so the narrow tail of one batch runs under the
        wide kernels of the next. Returns the list of pinned host records (dict of tensors, plus "counts") after one final
        synchronisation; a batch whose counts exceed the static capacities must be re-scored with score_host."""
        if hasattr(self.model, "ensure_inference_current"):
            self.model.ensure_inference_current()
        graphs = getattr(self.model, "_full_graphs", None)
        if not graphs:
            return [dict((k, torch.from_numpy(v)) for k, v in self.score_host(b).items()) for b in staged_batches]
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_slot_streams", None) is None or len(self._slot_streams) != len(graphs):
            self._slot_streams = [torch.cuda.Stream(self.device) for _ in graphs]
        for st in self._slot_streams:
            st.wait_stream(main)
        outs = []
        for i, b in enumerate(staged_batches):
            sl = i % len(graphs)
            with torch.cuda.stream(self._slot_streams[sl]):
                outs.append(self.fetch_async(self.model.full_graph_replay(b[0], b[1], slot=sl)))
        for st in self._slot_streams:
            main.wait_stream(st)
        main.synchronize()
        return outs

    def fetch_async(self, rec):
        """Queues the D2H copy of the record fields into pinned host tensors on the current stream (no sync)."""
        out = {}
        for k in RECORD_FIELDS + (("counts",) if "counts" in rec else ()):
            out[k] = torch.empty(rec[k].shape, dtype=rec[k].dtype, pin_memory=True)
            out[k].copy_(rec[k], non_blocking=True)
        return out

    # -- two-stream software pipeline ----------------------------------------------------------------------------
    def _geom_tensors(self, dev_batch, geom):
        yield dev_batch[0]
        yield dev_batch[1]
        yield geom["voxel_features"]
        yield geom["voxel_coords"]
        for key, d in geom["rulebooks"].items():
            if key == "_crb3d_cellmaps":          # cell -> row maps of the strided levels (spconv/pytorch/conv.py: _rulebook)
                for cm in d.values():
                    yield cm.ws
                continue
            for t in (d.nbr, d.nbr_t, d.out_indices, d.indices):
                if t is not None:
                    yield t

    def score_stream(self, batches, from_host=True, between_steps=None):
        """Generator over per-batch records. The geometry phase of batch i+1 (H2D copy, voxelize + MeanVFE, 8 rulebooks
        and their 5 host reads of counts) runs on a side stream while the feature phase of batch i (12 sparse convs, BEV
        stack, head, NMS, density, entropy) occupies the main stream, so the host reads never drain the GPU.
        batches: staged host batches (stage_host) when from_host else device batches (to_device)."""
        main = torch.cuda.current_stream(self.device)
        if getattr(self, "_side", None) is None:
            self._side = torch.cuda.Stream(self.device)
        side = self._side

        def launch_geometry(b):
            side.wait_stream(main) if launch_geometry.first else None
            launch_geometry.first = False
            with torch.cuda.stream(side):
                dev = self.to_device(b) if from_host else b
                geom = self.model.geometry(dev[0], dev[1], dev[1].numel() - 1)
                ev = torch.cuda.Event()
                ev.record(side)
            return dev, geom, ev

        launch_geometry.first = True
        it = iter(batches)
        try:
            nxt = launch_geometry(next(it))
        except StopIteration:
            return
        while nxt is not None:
            dev, geom, ev = nxt
            main.wait_event(ev)
            for t in self._geom_tensors(dev, geom):
                t.record_stream(main)
            rec = self.model.score_batch(dev[0], dev[1], dev[1].numel() - 1, dev[2], geom=geom)
            if between_steps is not None:
                between_steps()
            try:
                nxt = launch_geometry(next(it))      # host blocks on side-stream counts while `main` computes
            except StopIteration:
                nxt = None
            yield rec

    def record_tensor(self, rec, frame_ids):
        """Fixed-stride per-frame record (B, 3 + 2P) float32: [frame_id, num_boxes, entropy, labels..., density...]."""
        B = rec["entropy"].shape[0]
        out = torch.zeros((B, 3 + 2 * self.P), dtype=torch.float32, device=self.device)
        out[:, 0] = torch.as_tensor(frame_ids, dtype=torch.float32, device=self.device)
        out[:, 1] = rec["num_boxes"].float()
        out[:, 2] = rec["entropy"]
        out[:, 3:3 + self.P] = rec["labels"].float()
        out[:, 3 + self.P:] = rec["density"]
        return out

    def _slot_buffers(self, fg, n_slots):
        """Reusable pinned staging per graph copy: input points / offsets and the record fields coming back (a real pool is
        thousands of frames: nothing here grows with the pool size)."""
        key = (fg["cap"], fg["B"], n_slots, len(fg["caps"]))
        if getattr(self, "_pool_bufs", None) is not None and self._pool_bufs[0] == key:
            return self._pool_bufs[1]
        B, P = fg["B"], self.P
        bufs = []
        for _ in range(n_slots):
            bufs.append(dict(
                pts=torch.empty((fg["cap"], self.n_feat), dtype=torch.float32).pin_memory(),
                offs=torch.zeros((B + 1,), dtype=torch.int32).pin_memory(),
                entropy=torch.empty((B,), dtype=torch.float32).pin_memory(), num_boxes=torch.empty((B,), dtype=torch.int32).pin_memory(),
                labels=torch.empty((B, P), dtype=torch.int32).pin_memory(), density=torch.empty((B, P), dtype=torch.float32).pin_memory(),
                counts=torch.empty((len(fg["caps"]),), dtype=torch.int32).pin_memory(), event=torch.cuda.Event(), sel=None))
        self._pool_bufs = (key, bufs)
        return bufs

    def prepare_graphs(self, sample_frames, slots=2, margin=1.3, max_points_per_frame=None):
        """Captures the whole-step graphs for this scorer's batch size with static capacities measured on `sample_frames`
        (a list of >= batch_size host frames representative of the pool): row capacities = max count per level x margin,
        point capacity = the largest sample frame + 1024 unless given. score_pool() then streams full batches through them and
        re-scores any batch that exceeds a capacity on the eager path."""
        B = self.batch_size
        batches = [self.to_device(self.stage_host(sample_frames[s:s + B])) for s in range(0, len(sample_frames) - B + 1, B)]
        if not batches:
            raise ValueError("prepare_graphs needs at least batch_size sample frames")
        caps = self.model.calibrate_row_caps([(b[0], b[1], B) for b in batches], margin=margin)
        mx = max_points_per_frame or (max(len(f) for f in sample_frames) + 1024)
        self.model.enable_full_graph(B, max_points_per_frame=int(mx), slots=slots, row_caps=caps)
        return caps

    def score_pool(self, frames, frame_ids=None):
        """Scores this rank's shard of `frames` (all ranks pass the same list) and all-gathers the records.
        Returns a dict frame_id -> dict(entropy, labels (n,), density (n,)) identical on every rank.
        Full batches stream through the captured whole-step graphs (one copy per stream): frames are packed into a reusable
        pinned buffer of the copy, copied to the device, the graph replayed and the record copied back into pinned memory,
        while the host already packs the next batch; at most `slots` batches are in flight and every batch's row counts
        are checked against the static capacities as it completes. A partial last batch, a batch that does not fit the
        static buffers or one that exceeded a capacity is scored on the eager path (host-visible counts)."""
        world, rank = _world_rank()
        if hasattr(self.model, "ensure_inference_current"):   # retrained since the plan / graphs were built?
            self.model.ensure_inference_current()
        ids = list(range(len(frames))) if frame_ids is None else list(frame_ids)
        mine = shard_indices(len(frames), rank, world)
        sels = [mine[s:s + self.batch_size] for s in range(0, len(mine), self.batch_size)]
        P = self.P
        rows = np.zeros((len(mine), 3 + 2 * P), dtype=np.float32)
        row_of = {fi: r for r, fi in enumerate(mine)}
        graphs = getattr(self.model, "_full_graphs", None)
        fg = getattr(self.model, "_full_graph", None)
        eager = []

        def put(sel, h):
            for j, fi in enumerate(sel):
                r = row_of[fi]
                rows[r, 0], rows[r, 1], rows[r, 2] = fi, h["num_boxes"][j], h["entropy"][j]
                rows[r, 3:3 + P] = h["labels"][j]
                rows[r, 3 + P:] = h["density"][j]

        if graphs and fg is not None:
            caps = np.asarray(fg["caps"])
            bufs = self._slot_buffers(fg, len(graphs))
            main = torch.cuda.current_stream(self.device)
            if getattr(self, "_slot_streams", None) is None or len(self._slot_streams) != len(graphs):
                self._slot_streams = [torch.cuda.Stream(self.device) for _ in graphs]
            for st in self._slot_streams:
                st.wait_stream(main)

            def drain(buf):
                if buf["sel"] is None:
                    return
                buf["event"].synchronize()
                if bool((buf["counts"].numpy() <= caps).all()):
                    put(buf["sel"], {k: buf[k].numpy() for k in RECORD_FIELDS})
                else:
                    eager.append(buf["sel"])
                buf["sel"] = None

            i = 0
            for sel in sels:
                n = sum(len(frames[fi]) for fi in sel)
                mx = max(len(frames[fi]) for fi in sel)
                if len(sel) != fg["B"] or n > fg["cap"] or mx > fg["max_pts"]:
                    eager.append(sel)
                    continue
                sl = i % len(graphs)
                buf = bufs[sl]
                drain(buf)                                   # the copy's previous batch: record consumed, pinned buffers free
                pts_np, offs_np = buf["pts"].numpy(), buf["offs"].numpy()
                o = 0
                for j, fi in enumerate(sel):
                    f = frames[fi]
                    pts_np[o:o + len(f)] = f[:, -self.n_feat:]
                    o += len(f)
                    offs_np[j + 1] = o
                with torch.cuda.stream(self._slot_streams[sl]):
                    rec = self.model.full_graph_replay(buf["pts"][:n], buf["offs"], slot=sl)
                    for k in RECORD_FIELDS + ("counts",):
                        buf[k].copy_(rec[k], non_blocking=True)
                    buf["event"].record()
                buf["sel"] = sel
                i += 1
            for buf in bufs:
                drain(buf)
            for st in self._slot_streams:
                main.wait_stream(st)
        else:
            eager = list(sels)
        for sel in eager:
            put(sel, self._score_host_eager(self.stage_host([frames[fi] for fi in sel])))
        local = torch.from_numpy(rows).to(self.device)
        out = gather_records(local, len(frames), self.P, self.device)
        return {ids[k]: v for k, v in out.items()}

    def _score_host_eager(self, staged):
        """score_host on the dynamic path (host-visible counts), regardless of captured graphs."""
        dev = self.to_device(staged)
        geom = self.model.geometry(dev[0], dev[1], dev[1].numel() - 1)
        rec = self.model.score_batch(dev[0], dev[1], dev[1].numel() - 1, dev[2], geom=geom)
        out = {k: rec[k].to("cpu", non_blocking=True) for k in RECORD_FIELDS}
        torch.cuda.current_stream(self.device).synchronize()
        return {k: v.numpy() for k, v in out.items()}


def _world_rank():
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(), dist.get_rank()
    return 1, 0


def shard_indices(n_frames, rank, world):
    """frame i -> rank i mod W (the reference's eval DistributedSampler, pcdet/datasets/__init__.py:26-46)."""
    return list(range(rank, n_frames, world))


def gather_records(local, n_frames, P, device):
    """ONE all-gather of the fixed-stride records (row = [frame index, num_boxes, entropy, P labels, P densities]); rows
    are padded per rank to ceil(n/W) with frame index -1. Returns {frame index: dict(entropy, labels, density)}."""
    world, _ = _world_rank()
    per_rank = (n_frames + world - 1) // world
    pad = torch.full((per_rank, 3 + 2 * P), -1.0, dtype=torch.float32, device=device)
    pad[: local.shape[0]] = local
    if world > 1:
        gathered = torch.empty((world * per_rank, pad.shape[1]), dtype=torch.float32, device=device)
        dist.all_gather_into_tensor(gathered, pad)            # the single collective of the scoring path
    else:
        gathered = pad
    out = {}
    for row in gathered.cpu().numpy():
        if row[0] < 0:
            continue
        n = int(row[1])
        out[int(row[0])] = dict(entropy=float(row[2]), labels=row[3:3 + n].astype(np.int64),
                                density=row[3 + P:3 + P + n].astype(np.float32))
    return out
