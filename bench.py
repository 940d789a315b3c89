#!/usr/bin/env python
"""bench.py - frames/s of the SECOND forward + CRB stage-1 score on synthetic KITTI-shaped clouds (BASELINE.json metric).

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch B] [--impl reference]

Own arm: one step = one batch of B frames through crb3d.second.SECONDNet.score_batch (voxelize+MeanVFE -> 8 rulebooks ->
12 sparse convs -> dense -> BEV backbone -> anchor head -> score/top-k/decode -> batched rotated NMS -> points-in-boxes
density -> label entropy). `value` is timed with CUDA events per step (inputs resident in HBM, L2 flushed between steps),
`e2e` goes through the public PoolScorer.score_host call from pinned host memory with the D2H of the record inside the
timed region. Reference arm (--impl reference): the CPU restatement of the reference path (oracle/second_ref.py, kind
"port" - spconv is not installable here) on the host cores, one frame per step.
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
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=4)
    ap.add_argument("--impl", default="own", choices=["own", "reference"])
    ap.add_argument("--cpu-sample-frames", type=int, default=2)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of replaying the whole-step CUDA graph")
    ap.add_argument("--half-graph", action="store_true", help="round-1 mode: only the dense half in a CUDA graph, geometry pipelined on a side stream")
    ap.add_argument("--slots", type=int, default=4, help="independent copies of the whole-step graph replayed on alternating streams")
    ap.add_argument("--no-pipeline", action="store_true", help="(with --half-graph/--no-graph) one batch at a time on one stream")
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
def build_model(device):
    from crb3d import second
    torch.manual_seed(0)
    model = second.SECONDNet().eval().to_device(device)
    return model


def make_batches(batch):
    from crb3d import synth
    frames = [synth.make_frame(i) for i in range(N_DISTINCT_FRAMES)]
    return frames, [frames[s:s + batch] for s in range(0, N_DISTINCT_FRAMES - batch + 1, batch)]


def spconv_algorithmic_bytes(r):
    # 4*(sum_k P_k*C_in [gathered rows] + N_out*C_out [one write] + K*C_in*C_out [weights]) + 8*sum_k P_k
    return 4 * (r["pairs"] * r["cin"] + r["n_out"] * r["cout"] + r["K"] * r["cin"] * r["cout"]) + 8 * r["pairs"]


# ------------------------------------------------------------------------------------------------ reference arm
def run_reference(args, rank, world):
    if rank != 0:
        return
    from crb3d import head_ops, second, synth
    from oracle import second_ref
    torch.manual_seed(0)
    model = second.SECONDNet().eval()
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    # same head-bias calibration as the GPU arm needs a forward; use the CPU oracle logits of one frame
    anchors = head_ops.anchors_tensor(model.dense_head.spec)
    frames = [synth.make_frame(i) for i in range(max(args.steps + args.warmup, 1))]
    sd = {k: v.detach().clone() for k, v in model.state_dict().items()}
    second_ref.CALIBRATE_BN = True      # CPU twin of second.calibrate_batchnorm, on the first batch of the GPU arm
    try:
        second_ref.score_frames(sd, model.cfg, [synth.make_frame(i) for i in range(args.batch)], anchors)
    finally:
        second_ref.CALIBRATE_BN = False
    _calibrate_cpu(sd, model, frames[0], anchors)
    for i in range(args.warmup):
        second_ref.score_frames(sd, model.cfg, [frames[i]], anchors)
    t0 = time.perf_counter()
    for i in range(args.steps):
        second_ref.score_frames(sd, model.cfg, [frames[args.warmup + i]], anchors)
    dt = time.perf_counter() - t0
    fps = args.steps / dt
    line = {"impl": "reference", "metric": METRIC, "value": fps, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dt / max(args.steps, 1) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic (same generator and weight calibration as the own arm, done on the CPU)",
            "config": workload_config(args.batch, args.gpus),
            "cpu_baseline": {"value": fps, "unit": UNIT, "cores": cores, "kind": "port",
                             "sample": "%d steps x 1 frame (batch=1) of the same synthetic KITTI frames, torch threads=%d" % (args.steps, cores)},
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
    frames, batches = make_batches(args.batch)
    ps = scorer.PoolScorer(model, device, args.batch)
    staged = [ps.stage_host(b) for b in batches]
    resident = [ps.to_device(s) for s in staged]
    second.calibrate_batchnorm(model, resident[0][0], resident[0][1], args.batch)      # see the docstring: random init collapses
    second.calibrate_head_bias(model, resident[0][0], resident[0][1], args.batch, target_fraction=0.004)
    full_graph = not args.no_graph and not args.half_graph
    progress("calibrated; capturing graphs")
    if full_graph:          # the WHOLE step (voxelize .. entropy) as one CUDA graph, every count device-side
        model.enable_full_graph(args.batch, max_points_per_frame=max(s[2] for s in staged) + 1024, slots=max(1, args.slots))
        slot_streams = [torch.cuda.Stream(device) for _ in range(max(1, args.slots))]
    elif not args.no_graph:  # BEV backbone + head + post-processing (static shapes) as one CUDA graph
        model.enable_cuda_graph(args.batch, max_points_per_frame=max(s[2] for s in staged) + 1024)
    flush = torch.empty(160 << 20, dtype=torch.uint8, device=device)   # > 126 MB L2
    nb = len(resident)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(device)

    # pair counts of every sparse-conv launch of every distinct batch (outside any timed region)
    def dynamic_score(dev_batch):   # eager path with host-visible counts (the instrumented passes need per-launch hooks)
        geom = model.geometry(dev_batch[0], dev_batch[1], dev_batch[1].numel() - 1)
        return model.score_batch(dev_batch[0], dev_batch[1], dev_batch[1].numel() - 1, dev_batch[2], geom=geom)

    progress("graphs captured; counting pairs")
    pair_records = []
    for b in range(nb):
        ops.PROFILE = {"mode": "pairs", "records": []}
        dynamic_score(resident[b])
        pair_records.append(ops.PROFILE["records"])
    ops.PROFILE = None
    serial = args.no_pipeline or full_graph
    if serial:
        for i in range(args.warmup):
            ps.score_device(resident[i % nb])
    else:  # warm the side stream's allocator pool too: same code path as the timed region
        for _ in ps.score_stream((resident[i % nb] for i in range(max(args.warmup, 3))), from_host=False):
            pass
    progress("warm-up done")
    barrier()

    # ---- timed region 1: device-resident inputs ---------------------------------------------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    k0 = _lib.LAUNCHES["kernels"]
    rec = None
    fill = [0]

    def l2_flush():
        fill[0] += 1
        flush.fill_(fill[0] & 0xFF)

    barrier()
    ev0.record()
    if full_graph:      # consecutive batches alternate between the graph copies / streams; every step still flushes L2
        main = torch.cuda.current_stream(device)
        for st in slot_streams:
            st.wait_stream(main)
        for i in range(args.steps):
            sl = i % len(slot_streams)
            with torch.cuda.stream(slot_streams[sl]):
                rec = model.full_graph_replay(resident[i % nb][0], resident[i % nb][1], slot=sl)
                l2_flush()
        for st in slot_streams:
            main.wait_stream(st)
    elif serial:
        for i in range(args.steps):
            rec = ps.score_device(resident[i % nb])
            l2_flush()
    else:  # geometry of batch i+1 on a side stream under the feature phase of batch i (PoolScorer.score_stream)
        for rec in ps.score_stream((resident[i % nb] for i in range(args.steps)), from_host=False, between_steps=l2_flush):
            pass
    if world > 1:  # the single collective of the scoring path: all-gather of the (tiny) per-frame records
        local = ps.record_tensor(rec, list(range(args.batch)))
        gathered = torch.empty((world * local.shape[0], local.shape[1]), device=device)
        dist.all_gather_into_tensor(gathered, local)
    ev1.record()
    barrier()
    launches = _lib.LAUNCHES["kernels"] - k0
    total_ms = ev0.elapsed_time(ev1)
    progress("timed region 1 done (%.2f ms)" % total_ms)
    overflow = False
    if full_graph:
        overflow = not bool((rec["counts"].cpu().numpy() <= np.asarray(model._full_graph["caps"])).all())

    # ---- timed region 2: end to end through the public API (pinned host -> device -> host record) --------------
    for i in range(min(args.warmup, 2)):
        ps.score_host(staged[i % nb])
    barrier()
    e2e_t0 = time.perf_counter()
    outs = []
    if full_graph:      # the public throughput call: H2D into the static buffers, graph launch, D2H of the record, per batch
        outs = ps.score_host_stream([staged[i % nb] for i in range(args.steps)])
        out = {k: v.numpy() for k, v in outs[-1].items() if k != "counts"}
    elif args.no_pipeline:
        for i in range(args.steps):
            out = ps.score_host(staged[i % nb])
    else:
        for rec in ps.score_stream((staged[i % nb] for i in range(args.steps)), from_host=True):
            outs.append(ps.fetch_async(rec))            # D2H of the record into pinned memory, inside the timed region
        torch.cuda.synchronize(device)
        out = {k: v.numpy() for k, v in outs[-1].items()}
    torch.cuda.synchronize(device)
    e2e_ms = (time.perf_counter() - e2e_t0) * 1e3
    progress("timed region 2 (e2e) done (%.2f ms)" % e2e_ms)
    sampler.stop_flag = True
    sampler.join(timeout=2)
    h2d = int(np.mean([s[0].numel() * 4 + s[1].numel() * 4 for s in staged]))
    d2h = int(sum(v.nbytes for v in out.values()))

    # ---- instrumented pass (outside both timed regions): per-launch events of the heaviest kernels, eager path
    prof_records = {"mode": "time", "records": [], "conv2d": []}
    ops.PROFILE = prof_records
    for i in range(args.steps):
        dynamic_score(resident[i % nb])
        l2_flush()
    ops.PROFILE = None
    torch.cuda.synchronize(device)
    _lib.raise_if_device_error()
    progress("instrumented pass done")

    # max over ranks
    if world > 1:
        t = torch.tensor([total_ms, e2e_ms], device=device, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        total_ms, e2e_ms = float(t[0]), float(t[1])
    if rank != 0:
        return
    frames_total = args.steps * args.batch * world
    value = frames_total / (total_ms / 1e3)
    e2e_value = frames_total / (e2e_ms / 1e3)

    # ---- roofline: per-launch CUDA events of the two heaviest kernel families of this library --------------------
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    traffic = {}
    try:
        traffic = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
    except Exception:
        pass
    step_ms_avg = total_ms / max(args.steps, 1)
    # (a) sparse-conv forward, HBM bound: algorithmic bytes / time
    conv_total_ms = sum(a.elapsed_time(b) for a, b in prof_records["records"])
    alg_bytes = sum(sum(spconv_algorithmic_bytes(r) for r in pair_records[i % nb]) for i in range(args.steps))
    hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
    sp_ach = alg_bytes / (conv_total_ms / 1e3) / 1e9 if conv_total_ms > 0 else 0.0
    roof_sp = {"kernel": "spconv_fwd_tc / spconv_fwd_simt (12 sparse-conv launches per step, aggregated)", "bound": "hbm",
               "achieved": sp_ach, "peak": hbm_peak,
               "peak_source": "MEASURED_PEAKS.json hbm_gbs (of measured)" if peaks else "fallback 6650 GB/s (of fallback)",
               "unit": "GB/s", "frac": sp_ach / hbm_peak, "traffic": traffic.get("spconv_fwd_bytes_per_step"),
               "launches_per_step": len(pair_records[0]), "algorithmic_bytes_per_step": alg_bytes / max(args.steps, 1),
               "kernel_ms_per_step": conv_total_ms / max(args.steps, 1),
               "share_of_step": conv_total_ms / max(args.steps, 1) / step_ms_avg}
    # (b) bev_conv3x3_pair_tc, tensor bound: 2*9*C_in*C_out*pixels flops / time; TF32 peak = half the measured bf16 rate
    c2 = prof_records["conv2d"]
    c2_ms = sum(a.elapsed_time(b) for a, b, _ in c2)
    c2_flops = sum(f for _, _, f in c2)
    tf32_peak = float(peaks.get("bf16_tflops_sustained", 1400.0)) / 2.0
    c2_ach = c2_flops / (c2_ms / 1e3) / 1e12 if c2_ms > 0 else 0.0
    roof_c2 = {"kernel": "bev_conv3x3_pair_tc (halo-tile tcgen05 cta_group::2 TF32 conv, %d launches per step)" % (len(c2) // max(args.steps, 1)),
               "bound": "tensor", "achieved": c2_ach, "peak": tf32_peak,
               "peak_source": ("MEASURED_PEAKS.json bf16_tflops_sustained / 2 (TF32 = half the bf16 rate; of measured)" if peaks
                               else "fallback 1400/2 TFLOP/s (of fallback)"),
               "unit": "TFLOP/s", "frac": c2_ach / tf32_peak, "traffic": traffic.get("bev_conv3x3_bytes_per_launch"),
               "launches_per_step": len(c2) // max(args.steps, 1), "flops_per_step": c2_flops / max(args.steps, 1),
               "kernel_ms_per_step": c2_ms / max(args.steps, 1), "share_of_step": c2_ms / max(args.steps, 1) / step_ms_avg}
    roofline, roofline2 = (roof_c2, roof_sp) if c2_ms > conv_total_ms else (roof_sp, roof_c2)
    roofline["measured"] = roofline2["measured"] = ("CUDA events around every launch of the kernel on its launching stream during an "
                                                    "instrumented eager pass of the same %d steps (the headline region replays one CUDA "
                                                    "graph per step, which has no per-kernel events); L2 flushed between steps" % args.steps)

    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
            "ms_per_step": total_ms / max(args.steps, 1), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": ("f32 storage; tensor-core layers multiply in TF32 with fp32 accumulation (sparse convs, BEV convs, deblock/head "
                      "GEMMs; cuDNN TF32 where it is used is the reference's PyTorch default); --exact-fp32 switches the sparse "
                      "convs to fp32 FFMA") if not args.exact_fp32 else "f32 (sparse convs fp32 FFMA; BEV stack TF32)",
            "data": "synthetic (LiDAR-like clouds of SURVEY 8d; random-init SECOND weights with BatchNorm statistics taken from one "
                    "synthetic batch and the class-head bias calibrated so that ~1 % of the anchors pass SCORE_THRESH)",
            "config": workload_config(args.batch, world),
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h},
            "gpu_launches": launches, "clocks": sampler.summary(), "roofline": roofline, "roofline_secondary": roofline2}
    line["config"]["graph"] = ("whole step = one CUDA graph (device-side counts), %d copies replayed on alternating streams; static "
                               "capacities exceeded: %s" % (max(1, args.slots), overflow)) if full_graph \
        else ("dense half in a CUDA graph" if not args.no_graph else "eager")

    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(model, frames, args.cpu_sample_frames)
    print(json.dumps(line))


def cpu_baseline(model, frames, n_frames):
    """The oracle (kind 'port') timed on the host cores on a bounded sample of the same frames."""
    from crb3d import head_ops
    from oracle import second_ref
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    anchors = head_ops.anchors_tensor(model.dense_head.spec)
    sd = {k: v.detach().cpu() for k, v in model.state_dict().items()}
    second_ref.score_frames(sd, model.cfg, [frames[0]], anchors)            # warm-up (page-in, thread pool)
    t0 = time.perf_counter()
    for i in range(n_frames):
        second_ref.score_frames(sd, model.cfg, [frames[1 + i]], anchors)
    dt = time.perf_counter() - t0
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
