#!/usr/bin/env python
"""bench.py - frames/s of the SECOND forward + CRB stage-1 score on synthetic KITTI-shaped clouds (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Workload (own arm): CRB stage-1 scoring of an unlabelled pool of K*256 frames (default K=16: the 4096-frame pool of
BASELINE.json configs[3]) on the configs[1] shapes, the pool sharded frame i -> rank i mod W (strong scaling), ONE all-gather
of every frame's record at the end. One batch of 4 frames = one replay of the whole-step CUDA graph (voxelize+MeanVFE -> 8
rulebooks -> 12 sparse convs -> dense -> BEV backbone -> anchor head -> score/top-k/decode -> batched rotated NMS ->
points-in-boxes density -> label entropy); a step = 256 frames of the pool.
  value : CUDA events around the whole pool with the batches resident in HBM, an L2 flush after every replay, every record
          written into the rank's row block and the all-gather inside the region; max over ranks.
  e2e   : wall clock of the public call PoolScorer.score_pool(host frames) -> {frame: record}: host staging into pinned
          memory, H2D, replays, D2H of every record, the all-gather.
  extra : (N=1) exact-fp32 frames/s, a forward+backward step (configs[1]), Waymo-shaped batch 2 (configs[4]), per-stage table.
Reference arm (--impl reference): the CPU restatement of the reference path (oracle/second_ref.py, kind "port" - spconv is not
installable here) on all host cores; each step is ONE frame of the same pool (a bounded sample of the 256-frame step).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (os.path.join(ROOT, "crb-active-3ddet_b200"), ROOT):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "point-cloud frames/sec (fwd+CRB-score) SECOND-KITTI synthetic"
UNIT = "frames/s"
N_DISTINCT_FRAMES = 16


def progress(msg):
    """Per-phase progress on stderr: the tail of a killed run shows where it stopped."""
    sys.stderr.write("[bench rank %s %.1fs] %s\n" % (os.environ.get("RANK", "0"), time.perf_counter() - _T0, msg))
    sys.stderr.flush()


_T0 = time.perf_counter()


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=16, help="one step = 256 frames of the pool (default 16 steps = the 4096-frame pool of configs[3])")
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=16, help="frames per graph replay (the reference's eval batch, scripts/kitti/train_kitti_crb.sh:18)")
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the secondary lines (exact fp32, train step, Waymo)")
    ap.add_argument("--slots", type=int, default=2, help="independent copies of the whole-step graph replayed on alternating streams")
    ap.add_argument("--profile-replays", type=int, default=0, help="profiling aid (ncu launch list): after the warm-up run this many graph "
                    "replays on ONE stream and exit without timing anything")
    ap.add_argument("--exact-fp32", action="store_true", help="sparse convs on the exact-fp32 SIMT kernel instead of tcgen05 TF32")
    return ap.parse_args()


def workload_config(batch, n_gpus):
    return {"workload": "SECOND KITTI-synthetic (configs[1] shapes: ~20k pts/frame, 1408x1600x40 voxel grid, batch=%d per GPU), "
                        "metric path = forward + CRB stage-1 score (entropy + per-box density record)" % batch,
            "batch_per_gpu": batch, "global_batch": batch * n_gpus, "frames_distinct": N_DISTINCT_FRAMES,
            "l2": "160 MiB buffer (> the 126 MB L2) written inside every step of the timed region (L2 flush); per-step activations (>1 GB) also exceed L2",
            "parallelism": "frames sharded over ranks (dp%d), one all-gather of score records at the end" % n_gpus}


# ------------------------------------------------------------------------------------------------ clocks sampler
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.stop_flag = False

    def run(self):
        # NVML in-process (a few microseconds per sample); falls back to the nvidia-smi line of B200_PROFILING.md
        try:
            import pynvml
            pynvml.nvmlInit()
            h = pynvml.nvmlDeviceGetHandleByIndex(self.index)
            mx = pynvml.nvmlDeviceGetMaxClockInfo(h, pynvml.NVML_CLOCK_SM)
            R = pynvml
            bits = [(R.nvmlClocksEventReasonHwSlowdown, 0), (R.nvmlClocksEventReasonHwThermalSlowdown, 1),
                    (R.nvmlClocksEventReasonSwThermalSlowdown, 2), (R.nvmlClocksEventReasonSwPowerCap, 3)]
            while not self.stop_flag:
                sm = pynvml.nvmlDeviceGetClockInfo(h, pynvml.NVML_CLOCK_SM)
                try:
                    r = pynvml.nvmlDeviceGetCurrentClocksEventReasons(h)
                except Exception:
                    r = pynvml.nvmlDeviceGetCurrentClocksThrottleReasons(h)
                flags = ["Not Active"] * 4
                for bit, i in bits:
                    if r & bit:
                        flags[i] = "Active"
                self.samples.append([str(sm), str(mx)] + flags)
                time.sleep(0.02)
            return
        except Exception:
            pass
        q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
            "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        if not self.samples:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unavailable"]}
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = max((int(s[1]) for s in self.samples if s[1].isdigit()), default=None)
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": reasons, "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------ model / data
FRAMES_PER_STEP = 256          # one step = 256 frames of the pool (16 replays of 16 frames), shared out over the ranks


def build_model(device, cfg=None):
    from crb3d import second
    torch.manual_seed(0)
    model = (second.SECONDNet(cfg) if cfg is not None else second.SECONDNet()).eval().to_device(device)
    return model


def distinct_frames(n=N_DISTINCT_FRAMES, cfg=None):
    from crb3d import synth
    return [synth.make_frame(i) if cfg is None else synth.make_frame(i, cfg) for i in range(n)]


def workload_config(batch, n_gpus, steps):
    pool = steps * FRAMES_PER_STEP
    return {"workload": "CRB stage-1 scoring of a synthetic unlabelled pool, SECOND backbone (configs[3] workload on configs[1] shapes: "
                        "~20k pts/frame, 1408x1600x40 voxel grid, %d frames per graph replay = the reference's eval batch): forward + per-frame score record "
                        "(label entropy + per-box labels and point densities), frame i -> rank i mod W, ONE all-gather of "
                        "every record at the end (inside the timed region)" % batch,
            "pool_frames": pool, "frames_per_step": FRAMES_PER_STEP, "batch_per_replay": batch, "frames_distinct": N_DISTINCT_FRAMES,
            "l2": "160 MiB buffer (> the 126 MB L2) written after every replay of the timed region (L2 flush); a replay's "
                  "activations (>1 GB) also exceed L2",
            "parallelism": "pool sharded over %d rank(s) (strong scaling: the pool is fixed, %d frames per rank), one NCCL "
                           "all-gather of the fixed-stride records" % (n_gpus, pool // max(n_gpus, 1))}


def spconv_algorithmic_bytes(r):
    # 4*(sum_k P_k*C_in [gathered rows] + N_out*C_out [one write] + K*C_in*C_out [weights]) + 8*sum_k P_k
    return 4 * (r["pairs"] * r["cin"] + r["n_out"] * r["cout"] + r["K"] * r["cin"] * r["cout"]) + 8 * r["pairs"]


def calibrated_cpu_state(model_cfg_batch):
    """Random-init SECOND weights + the two synthetic-weight calibrations (BatchNorm statistics of one batch, class-head
    bias) done on the CPU with the oracle - the CPU twin of what the own arm does on the device."""
    from crb3d import head_ops, second, synth
    from oracle import second_ref
    torch.manual_seed(0)
    model = second.SECONDNet().eval()
    anchors = head_ops.anchors_tensor(model.dense_head.spec)
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    second_ref.CALIBRATE_BN = True
    try:
        second_ref.score_frames(sd, model.cfg, [synth.make_frame(i) for i in range(model_cfg_batch)], anchors)
    finally:
        second_ref.CALIBRATE_BN = False
    _calibrate_cpu(sd, model, synth.make_frame(0), anchors)
    return model, sd, anchors


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    """The reference's CPU implementation of the path (oracle port: spconv is not installable here) on all host cores.
    One step = a bounded sample of the own arm's step: ONE frame of the same pool (the own arm's step is 256 frames);
    value = frames/s, directly comparable."""
    if rank != 0:
        return
    from oracle import second_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    model, sd, anchors = calibrated_cpu_state(args.batch)
    frames = distinct_frames()
    for i in range(args.warmup):
        second_ref.score_frames(sd, model.cfg, [frames[i % len(frames)]], anchors)
    t0 = time.perf_counter()
    for i in range(args.steps):
        second_ref.score_frames(sd, model.cfg, [frames[i % len(frames)]], anchors)     # pool frame i, as the own arm
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generator and weight calibration as the own arm, done on the CPU)",
            "config": workload_config(args.batch, args.gpus, args.steps),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "each step = ONE frame of the same pool (a bounded sample of the own arm's 256-frame step), batch=1: "
                                       "%d frames, torch threads=%d" % (args.steps, cores)},
            "e2e": {"value": fps, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line))


def _calibrate_cpu(sd, model, frame, anchors, target_fraction=0.004):
    """CPU twin of crb3d.second.calibrate_head_bias (synthetic weights need a bias that lets boxes clear SCORE_THRESH)."""
    from oracle import second_ref
    col = {}
    second_ref.score_frames(sd, model.cfg, [frame], anchors, collect=col)
    n_loc, nc = model.dense_head.n_loc, model.num_class
    logits = col["cls_preds"].reshape(-1, n_loc * nc)
    thr = float(np.log(model.cfg["score_thresh"] / (1 - model.cfg["score_thresh"])))
    mean, std = logits.mean(0), logits.std(0).clamp_min(1e-6)
    zs = ((logits - mean) / std).flatten()
    z = torch.quantile(zs[:: max(1, zs.numel() // 2000000)], 1.0 - target_fraction)
    w = sd["dense_head.conv_cls.weight"].view(n_loc * nc, -1)
    b = sd["dense_head.conv_cls.bias"].view(n_loc * nc)
    w /= std.view(-1, 1)
    b.copy_((b - mean) / std - z + thr)


# ------------------------------------------------------------------------------------------------ own arm
def run_own(args, rank, world, local_rank):
    import torch.distributed as dist
    from crb3d import _lib, ops, scorer, second
    _lib.load()  # fail loudly if the native library is missing
    device = torch.device("cuda", local_rank)
    torch.cuda.set_device(device)
    _lib.init_device(local_rank)
    progress("building the model")
    model = build_model(device)
    model.prepare_inference(fold_bev_bn=True, spconv_tf32=not args.exact_fp32)
    B = args.batch
    frames = distinct_frames()
    ps = scorer.PoolScorer(model, device, B)
    cal = ps.to_device(ps.stage_host([frames[i % len(frames)] for i in range(B)]))     # --batch may exceed the 16 distinct clouds
    second.calibrate_batchnorm(model, cal[0], cal[1], B)      # see the docstring: random init collapses
    second.calibrate_head_bias(model, cal[0], cal[1], B, target_fraction=0.004)
    progress("calibrated; capturing graphs")
    slots = max(1, args.slots)
    max_pts = max(len(f) for f in frames)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2

    # ---- the pool: steps * 256 frames (cycling over the 16 distinct clouds), frame i -> rank i mod W
    pool_n = args.steps * FRAMES_PER_STEP
    pool = [frames[i % len(frames)] for i in range(pool_n)]
    mine = scorer.shard_indices(pool_n, rank, world)
    sels = [mine[s:s + B] for s in range(0, len(mine) - B + 1, B)]           # full batches (pool_n/W is a multiple of 4 here)
    # device-resident copies of the distinct batches this rank will replay (inputs resident in HBM for `value`)
    resident = {}
    keys = [tuple(i % len(frames) for i in sel) for sel in sels]
    for sel, key in zip(sels, keys):
        if key not in resident:
            resident[key] = ps.to_device(ps.stage_host([pool[i] for i in sel]))
    seq = [resident[k] for k in keys]
    # static row capacities of the captured step: measured on this rank's distinct batches (+30 %); a batch that exceeds one is
    # detected from the returned counts (checked for EVERY timed replay below) and would be re-scored on the eager path
    row_caps = model.calibrate_row_caps([(db[0], db[1], B) for db in resident.values()], margin=1.3)
    model.enable_full_graph(B, max_points_per_frame=max_pts + 1024, slots=slots, row_caps=row_caps)
    slot_streams = [torch.cuda.Stream(device) for _ in range(slots)]
    caps = np.asarray(model._full_graph["caps"])
    progress("graphs captured (row capacities %s); %d frames in the pool, %d replays on this rank" % (list(caps), pool_n, len(seq)))

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    def dynamic_score(dev_batch):   # eager path with host-visible counts (the instrumented passes need per-launch hooks)
        geom = model.geometry(dev_batch[0], dev_batch[1], dev_batch[1].numel() - 1)
        return model.score_batch(dev_batch[0], dev_batch[1], dev_batch[1].numel() - 1, dev_batch[2], geom=geom)

    pair_records = {}
    for key, db in resident.items():     # pair counts of every sparse-conv launch of every distinct batch (outside any timed region)
        ops.PROFILE = {"mode": "pairs", "records": []}
        dynamic_score(db)
        pair_records[key] = ops.PROFILE["records"]
    ops.PROFILE = None
    for i in range(max(args.warmup, 3)):
        for sl in range(slots):
            with torch.cuda.stream(slot_streams[sl]):
                model.full_graph_replay(seq[(i * slots + sl) % len(seq)][0], seq[(i * slots + sl) % len(seq)][1], slot=sl)
    torch.cuda.synchronize(device)
    progress("warm-up done")
    if args.profile_replays > 0:     # the launch list of exactly the kernels one replay of the timed region runs (never a bench value)
        torch.cuda.synchronize(device)
        torch.cuda.profiler.start()          # `ncu --profile-from-start off` captures exactly these replays
        for i in range(args.profile_replays):
            model.full_graph_replay(seq[i % len(seq)][0], seq[i % len(seq)][1], slot=0)
            flush.fill_(i & 0xFF)
        torch.cuda.synchronize(device)
        torch.cuda.profiler.stop()
        progress("profile replays done")
        return

    # ---- timed region 1 (`value`): inputs resident in HBM; every replay's record lands in a device row block; one all-gather
    sampler = ClockSampler(local_rank)
    sampler.start()
    P = ps.P
    per_rank = (pool_n + world - 1) // world
    rows = torch.full((per_rank, 3 + 2 * P), -1.0, device=device)
    counts_all = torch.zeros((len(seq), len(caps)), dtype=torch.int32, device=device)
    fid_dev = [torch.tensor(sel, dtype=torch.float32, device=device) for sel in sels]
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = _lib.LAUNCHES["kernels"]
    fill = [0]

    def l2_flush():
        fill[0] += 1
        flush.fill_(fill[0] & 0xFF)

    barrier()
    main = torch.cuda.current_stream(device)
    ev0.record()
    for st in slot_streams:
        st.wait_stream(main)
    for i, db in enumerate(seq):
        sl = i % slots
        with torch.cuda.stream(slot_streams[sl]):
            rec = model.full_graph_replay(db[0], db[1], slot=sl)
            r0 = i * B
            rows[r0:r0 + B, 0] = fid_dev[i]
            rows[r0:r0 + B, 1] = rec["num_boxes"]
            rows[r0:r0 + B, 2] = rec["entropy"]
            rows[r0:r0 + B, 3:3 + P] = rec["labels"]
            rows[r0:r0 + B, 3 + P:] = rec["density"]
            counts_all[i] = rec["counts"]
            l2_flush()
    for st in slot_streams:
        main.wait_stream(st)
    if world > 1:   # the single collective of the scoring path: every rank's records
        gathered = torch.empty((world * per_rank, 3 + 2 * P), device=device)
        dist.all_gather_into_tensor(gathered, rows)
    ev1.record()
    barrier()
    launches = _lib.LAUNCHES["kernels"] - k0
    total_ms = ev0.elapsed_time(ev1)
    overflow = not bool((counts_all.cpu().numpy() <= caps[None, :]).all())      # EVERY timed replay is checked
    progress("timed region 1 done (%.2f ms, %d replays, overflow=%s)" % (total_ms, len(seq), overflow))
    _lib.raise_if_device_error()

    # ---- timed region 2 (`e2e`): the public call - host frames in, the gathered records of the WHOLE pool out
    ps.score_pool(pool[: world * B * slots * 2])          # warm the pinned staging buffers / streams
    barrier()
    e2e_t0 = time.perf_counter()
    recs = ps.score_pool(pool)
    torch.cuda.synchronize(device)
    if world > 1:
        dist.barrier()
    e2e_ms = (time.perf_counter() - e2e_t0) * 1e3
    assert len(recs) == pool_n
    sampler.stop_flag = True
    sampler.join(timeout=2)
    progress("timed region 2 (e2e) done (%.2f ms)" % e2e_ms)
    frames_per_rank_step = len(mine) / max(args.steps, 1)
    h2d = int(np.mean([len(f) for f in frames]) * 4 * 4 * frames_per_rank_step + (B + 1) * 4 * frames_per_rank_step / B)
    d2h = int(frames_per_rank_step / B * (B * 4 * 2 + B * P * 4 * 2 + len(caps) * 4))

    # ---- instrumented pass (outside both timed regions): per-launch events of the heaviest kernels + per-stage events
    n_inst = min(len(seq), 16)
    prof_records = {"mode": "time", "records": [], "conv2d": []}
    ops.PROFILE = prof_records
    for i in range(n_inst):
        dynamic_score(seq[i])
        l2_flush()
    ops.PROFILE = None
    torch.cuda.synchronize(device)
    stage_ms = instrumented_stages(model, seq[:4], ops)
    _lib.raise_if_device_error()
    progress("instrumented pass done")

    extra = {}
    if world == 1 and not args.no_extra:
        extra = extra_measurements(args, model, ps, frames, device, l2_flush, progress)

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    value = pool_n / (total_ms / 1e3)
    e2e_value = pool_n / (e2e_ms / 1e3)

    # ---- roofline: per-launch CUDA events of the two heaviest kernel families of this library --------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    replay_ms = total_ms / max(len(seq), 1)
    # (a) sparse-conv forward, HBM bound: algorithmic bytes / time
    conv_total_ms = sum(a.elapsed_time(b) for a, b in prof_records["records"])
    alg_bytes = sum(sum(spconv_algorithmic_bytes(r) for r in pair_records[keys[i]]) for i in range(n_inst))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sp_ach = alg_bytes / (conv_total_ms / 1e3) / 1e9 if conv_total_ms > 0 else 0.0
    roof_sp = {"kernel": "spconv_fwd_tc / spconv_fwd_simt (12 sparse-conv launches per replay, aggregated)", "bound": "hbm",
               "achieved": sp_ach, "peak": hbm_peak,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
               "unit": "GB/s", "frac": sp_ach / hbm_peak, "traffic": None,
               "launches_per_replay": len(pair_records[keys[0]]), "algorithmic_bytes_per_replay": alg_bytes / max(n_inst, 1),
               "kernel_ms_per_replay": conv_total_ms / max(n_inst, 1),
               "share_of_replay": conv_total_ms / max(n_inst, 1) / replay_ms}
    # (b) the tcgen05 BEV 3x3 convs, tensor bound: 2*9*C_in*C_out*pixels flops / time; TF32 peak = half the measured bf16 rate.
    # The kernel is event-timed alone in an eager pass (idle gaps between launches): the BURST figure is the fair denominator.
    c2 = prof_records["conv2d"]
    c2_ms = sum(a.elapsed_time(b) for a, b, _ in c2)
    c2_flops = sum((f() if callable(f) else f) for _, _, f in c2)      # sparse-tile layers report the tiles they computed
    tf32_peak = float(peaks.get("bf16_tflops", 1600.0)) / 2.0
    c2_ach = c2_flops / (c2_ms / 1e3) / 1e12 if c2_ms > 0 else 0.0
    roof_c2 = {"kernel": "bev_conv3x3_pair_tc (halo-tile tcgen05 cta_group::2 TF32 convs, %d launches per replay; block 1 in sparse-tile mode: "
                         "flops = the tiles that went through the tensor cores, time includes the constant-tile fill)" % (len(c2) // max(n_inst, 1)),
               "bound": "tensor", "achieved": c2_ach, "peak": tf32_peak,
               "peak_source": ("MEASURED_PEAKS.json bf16_tflops (burst) / 2 (TF32 = half the bf16 rate; of measured)" if peaks
                               else "fallback 1600/2 TFLOP/s (of fallback)"),
               "unit": "TFLOP/s", "frac": c2_ach / tf32_peak, "traffic": None,
               "launches_per_replay": len(c2) // max(n_inst, 1), "flops_per_replay": c2_flops / max(n_inst, 1),
               "kernel_ms_per_replay": c2_ms / max(n_inst, 1), "share_of_replay": c2_ms / max(n_inst, 1) / replay_ms}
    # DRAM bytes per launch from the committed `ncu --set full` capture of this command's replay (profiles/traffic.json); only for the
    # configuration the capture was made at, otherwise null
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        if B == 16 and not args.exact_fp32:
            roof_c2["traffic"], roof_sp["traffic"] = tj.get("bev_conv_bytes_per_launch"), tj.get("spconv_bytes_per_launch")
            roof_c2["traffic_source"] = roof_sp["traffic_source"] = "bytes per launch, " + tj.get("_source", "profiles/traffic.json")
    except Exception:
        pass
    roofline, roofline2 = (roof_c2, roof_sp) if c2_ms > conv_total_ms else (roof_sp, roof_c2)
    roofline["measured"] = roofline2["measured"] = ("CUDA events around every launch of the kernel on its launching stream during an "
                                                    "instrumented eager pass over %d batches of the pool (the headline region replays one CUDA "
                                                    "graph per batch, which has no per-kernel events); L2 flushed between batches" % n_inst)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "strong", "vs_baseline": None,
            "dtype": ("f32 storage; tensor-core layers multiply in TF32 with fp32 accumulation (sparse convs, BEV convs, deblock/head "
                      "GEMMs); --exact-fp32 switches the sparse convs to fp32 FFMA") if not args.exact_fp32 else "f32 (sparse convs fp32 FFMA; BEV stack TF32)",
            "data": "synthetic (LiDAR-like clouds of SURVEY 8d; random-init SECOND weights with BatchNorm statistics taken from one "
                    "synthetic batch and the class-head bias calibrated so that ~1 % of the anchors pass SCORE_THRESH)",
            "config": workload_config(B, world, args.steps),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "path": "PoolScorer.score_pool(host frames) -> {frame: record} incl. host staging, H2D, graph replays, D2H, all-gather"},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "roofline_secondary": roofline2,
            "stages": stage_ms}
    line["config"]["graph"] = ("one batch = one CUDA graph (device-side counts), %d copies replayed on %d streams; static capacities "
                               "exceeded in any timed replay: %s" % (slots, slots, overflow))
    if extra:
        line["extra"] = extra
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(model, frames, args.cpu_sample_frames, line["stages"])
    print(json.dumps(line))


def instrumented_stages(model, batches, ops):
    """SURVEY 8(d) per-stage table, device side: CUDA events around each stage of the eager path (ms per batch)."""
    dev = batches[0][0].device
    names = ["voxelize", "rulebooks", "sparse_backbone", "dense", "bev_head", "post", "density_entropy"]
    acc = dict((n, 0.0) for n in names)
    cfg = model.cfg
    for pts, offs, mx in batches:
        B = offs.numel() - 1
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(len(names) + 1)]
        with torch.no_grad():
            ev[0].record()
            vox = model.voxelize(pts, offs, B)
            ev[1].record()
            books = model.backbone_3d.build_rulebooks(vox["coords"], B)
            ev[2].record()
            bd = model.backbone_3d(dict(batch_size=B, voxel_features=vox["mean"], voxel_coords=vox["coords"], rulebooks=books))
            ev[3].record()
            bd = model.map_to_bev_module(bd)
            ev[4].record()
            bd = model.dense_head(model.backbone_2d(bd))
            ev[5].record()
            from crb3d import head_ops
            A = model.dense_head.num_anchors
            k = min(cfg["nms_pre_maxsize"], A)
            score, label, top_scores, top_idx, counts = head_ops.anchor_head_scores_topk(bd["cls_preds"], model.num_class, B, cfg["score_thresh"], k)
            boxes = head_ops.anchor_decode_select(bd["box_preds"], bd["dir_cls_preds"], top_idx, model.dense_head.spec, A)
            keep, num = ops.nms_batched(boxes, counts, cfg["nms_thresh"], rotated=True, max_keep=cfg["nms_post_maxsize"])
            final_boxes = head_ops.gather_rows(boxes, keep, num, 0.0)
            ev[6].record()
            P = cfg["nms_post_maxsize"]
            box_begin = torch.arange(B, device=dev, dtype=torch.int32) * P
            ops.points_in_boxes_ranges(pts, offs[:-1], offs[1:], mx, final_boxes, box_begin, box_begin + num)
            ev[7].record()
        torch.cuda.synchronize(dev)
        for i, n in enumerate(names):
            acc[n] += ev[i].elapsed_time(ev[i + 1])
    nb = int(batches[0][1].numel() - 1)
    return {"unit": "ms per batch of %d frames (eager launches, one stream, CUDA events)" % nb, "gpu": dict((n, acc[n] / len(batches)) for n in names)}


def extra_measurements(args, model, ps, frames, device, l2_flush, progress):
    """Secondary lines (N=1 only, each a few seconds): exact-fp32 frames/s, a forward+backward step (configs[1] names it),
    the Waymo-shaped configuration (configs[4]) and the 20-replay figure of round 1."""
    from crb3d import ops, scorer, second
    out = {}
    B = args.batch
    slots = max(1, args.slots)
    batches = [ps.to_device(ps.stage_host([frames[(s + i) % len(frames)] for i in range(B)])) for s in range(0, max(len(frames) - B, 0) + 1, B)]
    b4 = ps.to_device(ps.stage_host(frames[:4]))
    row_caps = model._graph_cfg[4]

    def replay_rate(n, streams):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        main = torch.cuda.current_stream(device)
        torch.cuda.synchronize(device)
        ev0.record()
        for st in streams:
            st.wait_stream(main)
        for i in range(n):
            with torch.cuda.stream(streams[i % len(streams)]):
                model.full_graph_replay(batches[i % len(batches)][0], batches[i % len(batches)][1], slot=i % len(streams))
                l2_flush()
        for st in streams:
            main.wait_stream(st)
        ev1.record()
        torch.cuda.synchronize(device)
        return n * B / (ev0.elapsed_time(ev1) / 1e3)

    streams = [torch.cuda.Stream(device) for _ in range(slots)]
    out["replays_20_frames_per_s"] = replay_rate(20, streams)
    # exact fp32 sparse convs (the parity configuration: tests/test_gpu_second.py::test_pool_stage1_indices_and_ranking_exact)
    try:
        old = ops.SPCONV_TF32
        model.prepare_inference(fold_bev_bn=True, spconv_tf32=False)
        model.enable_full_graph(B, max_points_per_frame=max(len(f) for f in frames) + 1024, slots=slots, row_caps=row_caps)
        replay_rate(8, streams)
        out["exact_fp32_sparse_convs_frames_per_s"] = replay_rate(64, streams)
    finally:
        model.prepare_inference(fold_bev_bn=True, spconv_tf32=old)
        model.enable_full_graph(B, max_points_per_frame=max(len(f) for f in frames) + 1024, slots=slots, row_caps=row_caps)
    progress("extra: exact fp32 done")
    # forward + backward (configs[1] "fwd+bwd batch 4"): voxelize, rulebooks, every layer's forward and backward, target assignment,
    # the three anchor-head losses (anchor_head_template.py:101-229 on csrc/train_ops.cu) and an SGD step
    try:
        from crb3d import synth
        tm = build_model(device)
        tm.train()
        opt = torch.optim.SGD(tm.parameters(), lr=1e-4)
        pts, offs, _ = b4
        gts = [synth.make_frame(i, return_boxes=True)[1] for i in range(4)]
        gt = np.zeros((4, max(len(g) for g in gts), 8), np.float32)
        for i, g in enumerate(gts):
            gt[i, :len(g)] = g
        gt = torch.from_numpy(gt).to(device)

        def train_step():
            opt.zero_grad(set_to_none=True)
            tm.forward_features(pts, offs, 4, gt_boxes=gt)
            loss, _ = tm.dense_head.get_loss()
            loss.backward()
            opt.step()
        old = ops.SPCONV_TF32
        ops.SPCONV_TF32 = not args.exact_fp32
        for _ in range(3):
            train_step()
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize(device)
        ev0.record()
        for _ in range(10):
            train_step()
        ev1.record()
        torch.cuda.synchronize(device)
        ops.SPCONV_TF32 = old
        ms = ev0.elapsed_time(ev1) / 10
        out["train_step_fwd_bwd_batch4"] = {"ms_per_step": ms, "frames_per_s": 4 / (ms / 1e3),
                                            "note": "voxelize + 8 rulebooks + 12 sparse convs fwd/dX/dW + BEV stack fwd/bwd (cuDNN autograd) + "
                                                    "axis-aligned target assignment + focal / smooth-L1 / direction losses (csrc/train_ops.cu) + SGD"}
        del tm, opt
    except Exception as e:  # pragma: no cover
        out["train_step_fwd_bwd_batch4"] = {"error": str(e).splitlines()[0][:200]}
    progress("extra: train step done")
    # Waymo-shaped clouds (configs[4] shapes: ~160k pts, 1504x1504x40 grid, 5 features), batch 2, forward + stage-1 score
    try:
        from crb3d import synth
        wm = build_model(device, second.WAYMO_SECOND_CFG)
        wm.prepare_inference(fold_bev_bn=True, spconv_tf32=not args.exact_fp32)
        wf = [synth.make_frame(i, synth.WAYMO) for i in range(4)]
        wps = scorer.PoolScorer(wm, device, 2)
        wb = [wps.to_device(wps.stage_host(wf[s:s + 2])) for s in (0, 2)]
        second.calibrate_batchnorm(wm, wb[0][0], wb[0][1], 2)
        second.calibrate_head_bias(wm, wb[0][0], wb[0][1], 2, target_fraction=0.004)
        wm.enable_full_graph(2, max_points_per_frame=max(len(f) for f in wf) + 1024, slots=2,
                             row_caps=wm.calibrate_row_caps([(b[0], b[1], 2) for b in wb], margin=1.3))
        ws = [torch.cuda.Stream(device) for _ in range(2)]

        def wrate(n):
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            main = torch.cuda.current_stream(device)
            torch.cuda.synchronize(device)
            ev0.record()
            for st in ws:
                st.wait_stream(main)
            rec = None
            for i in range(n):
                with torch.cuda.stream(ws[i % 2]):
                    rec = wm.full_graph_replay(wb[i % 2][0], wb[i % 2][1], slot=i % 2)
                    l2_flush()
            for st in ws:
                main.wait_stream(st)
            ev1.record()
            torch.cuda.synchronize(device)
            ok = bool((rec["counts"].cpu().numpy() <= np.asarray(wm._full_graph["caps"])).all())
            return n * 2 / (ev0.elapsed_time(ev1) / 1e3), ok
        wrate(4)
        fps, ok = wrate(24)
        out["waymo_synthetic_batch2"] = {"frames_per_s": fps, "points_per_frame": int(np.mean([len(f) for f in wf])),
                                         "capacities_ok": ok, "note": "SECOND Waymo-synthetic forward + stage-1 score, 2 graph copies, 1 GPU"}
        del wm
    except Exception as e:  # pragma: no cover
        out["waymo_synthetic_batch2"] = {"error": str(e).splitlines()[0][:200]}
    progress("extra: waymo done")
    torch.cuda.empty_cache()
    return out


def cpu_baseline(model, frames, n_frames, stages=None):
    """The oracle (kind 'port') timed on the host cores on a bounded sample of the same frames, with its per-stage times
    (the CPU column of the SURVEY 8(d) stage table)."""
    from crb3d import head_ops
    from oracle import second_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    anchors = head_ops.anchors_tensor(model.dense_head.spec)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    second_ref.score_frames(sd, model.cfg, [frames[0]], anchors)            # warm-up (page-in, thread pool)
    timers = {}
    t0 = time.perf_counter()
    for i in range(n_frames):
        second_ref.score_frames(sd, model.cfg, [frames[1 + i]], anchors, timers=timers)
    dt = time.perf_counter() - t0
    if stages is not None:
        stages["cpu"] = dict((k, v / n_frames * 1e3) for k, v in timers.items())
        stages["cpu_unit"] = "ms per frame (oracle port, %d torch threads; sparse_backbone includes its rulebooks)" % cores
    return {"value": n_frames / dt, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%d frames, batch=1, same synthetic KITTI frames and weights (oracle/second_ref.py, torch threads=%d)" % (n_frames, cores)}


def main():
    args = parse()
    # stdout carries exactly ONE JSON line: libraries that print to fd 1 (NCCL's version banner, cuDNN logs) are sent to
    # stderr for the whole run and the result line is written to the saved descriptor at the end
    real_stdout = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    sys.stdout = real_stdout
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        import datetime
        # a rank that stops answering is reported after 2 minutes (default: 10), with the flight recorder's last collectives
        os.environ.setdefault("TORCH_NCCL_TRACE_BUFFER_SIZE", "2000")
        os.environ.setdefault("TORCH_NCCL_DUMP_ON_TIMEOUT", "1")
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank), timeout=datetime.timedelta(seconds=120))
    try:
        run_own(args, rank, world, local_rank)
    except Exception:
        from crb3d import _lib
        progress("FAILED; device diagnostics record: %r" % (_lib.last_device_error(),))
        raise
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
