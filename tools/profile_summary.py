#!/usr/bin/env python
"""Turns gpurun_out/ ncu outputs into the text summaries committed under profiles/.

  python tools/profile_summary.py launches gpurun_out/launches.csv N_STEPS > profiles/rNN_launches.txt
  python tools/profile_summary.py ncu gpurun_out/prof.ncu-rep > profiles/rNN_kernel.txt
"""
import collections
import csv
import re
import subprocess
import sys


def launches(path, steps):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = [r for r in csv.DictReader(lines) if r.get("Metric Name") == "gpu__time_duration.sum"]

    def us(r):
        v = float(r["Metric Value"].replace(",", ""))
        return v / 1000 if r["Metric Unit"] == "ns" else (v * 1000 if r["Metric Unit"] == "ms" else v)

    def nm(r):
        n = r["Kernel Name"].replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("pair::", "")
        return re.sub(r"\(.*", "", n).replace("void ", "")[:72]

    idx = [i for i, r in enumerate(rows) if nm(r).startswith("vox_insert")]
    start = idx[-steps] - 1
    sel = rows[start:]
    agg = collections.OrderedDict()
    total = 0.0
    for r in sel:
        a = agg.setdefault(nm(r), [0, 0.0])
        a[0] += 1
        a[1] += us(r)
        total += us(r)
    mine = ("vox_", "scan_", "fill_kernel", "subm_", "sparse_", "bitmap_", "pair_", "spconv_", "dense_kernel", "head_", "gather_rows",
            "nms_", "pairwise_kernel", "points_in_boxes", "box_density", "label_entropy", "roiaware", "ball_query", "group_points",
            "fps_kernel", "three_", "kde_", "sqdist", "bev_", "topk_", "round_tf32", "sat_", "tile_", "occ_scatter", "bitmap_", "fps_cluster",
            "sa_group", "fc_", "assign_", "head_loss", "count_pos", "range_flags", "collate_", "ff_", "row_norms", "voxel_query",
            "roipoint", "local_neighbors", "vector_pool", "gather_counts")
    own = sum(v for k, (c, v) in agg.items() if k.startswith(mine))
    print("# ncu launch list (gpu__time_duration.sum, --clock-control none; cold-cache, serialised: compare SHARES)")
    print("# last %d bench steps: %d launches, %.1f us per step; kernels of libcrb3d_sm100: %.1f%% of the time" %
          (steps, len(sel), total / steps, 100 * own / total))
    print("%12s %10s %7s  %s" % ("us/step", "calls/step", "share", "kernel"))
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print("%12.1f %10.1f %6.1f%%  %s%s" % (v / steps, c / steps, 100 * v / total, "* " if k.startswith(mine) else "  ", k))


def ncu(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    want = ["Kernel Name", "launch__grid_size", "launch__block_size", "launch__registers_per_thread", "gpu__time_duration.sum",
            "dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
            "lts__t_bytes.sum", "lts__t_sector_hit_rate.pct", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
            "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_lsu.sum", "smsp__cycles_active.avg", "launch__occupancy_limit_shared_mem",
            "launch__occupancy_limit_registers", "smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio",
            "smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio"]
    idx = [(w, hdr.index(w)) for w in want if w in hdr]
    print("# ncu --set full --clock-control none (one row per captured launch)")
    for r in rows[2:]:
        print("-" * 100)
        for w, i in idx:
            print("%-72s %s %s" % (w, r[i][:90], units[i]))


if __name__ == "__main__":
    if sys.argv[1] == "launches":
        launches(sys.argv[2], int(sys.argv[3]))
    else:
        ncu(sys.argv[2])
