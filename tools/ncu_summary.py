#!/usr/bin/env python
"""Condense an .ncu-rep (ncu --set full) into a small CSV of the metrics the roofline discussion uses.
usage: python tools/ncu_summary.py gpurun_out/prof.ncu-rep profiles/<name>.csv"""
import csv, io, subprocess, sys
KEEP = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "l1tex__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "launch__registers_per_thread", "launch__grid_size", "launch__block_size",
        "launch__shared_mem_per_block_dynamic"]
raw = subprocess.run(["ncu", "-i", sys.argv[1], "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
idx = [hdr.index(k) for k in KEEP if k in hdr]
with open(sys.argv[2], "w", newline="") as f:
    w = csv.writer(f)
    for r in rows:
        w.writerow([r[i] for i in idx])
print(open(sys.argv[2]).read())

# optional 3rd/4th (5th: commit the capture was taken at) argument: frames per launch of the captured command and the JSON bench.py reads for roofline.traffic
# (dram__bytes_read.sum + dram__bytes_write.sum of each kernel, per frame, averaged over its captured launches)
if len(sys.argv) >= 5:
    import json, re
    frames, out = int(sys.argv[3]), sys.argv[4]
    stage = {"k_order_winners": "order_winners", "k_order_claim": "order_winners", "k_order_scatter": "order_scatter",
             "k_order_fill": "order_scatter", "k_ground_mark": "ground_mark", "k_seg_build": "sector_mean", "k_seg_fold": "sector_mean",
             "k_sector_mean": "sector_mean", "k_finalize_bin": "finalize_bin_scatter"}
    col = {k: hdr.index(k) for k in ("Kernel Name", "dram__bytes_read.sum", "dram__bytes_write.sum")}
    units = rows[1]
    scale = {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}
    acc, cnt = {}, {}
    for r in rows[2:]:
        m = re.search(r"(k_[a-z_]+)", r[col["Kernel Name"]])
        if not m or m.group(1) not in stage:
            continue
        b = sum(float(r[col[c]]) * scale[units[col[c]]] for c in ("dram__bytes_read.sum", "dram__bytes_write.sum"))
        acc.setdefault(m.group(1), 0.0); cnt.setdefault(m.group(1), 0)
        acc[m.group(1)] += b; cnt[m.group(1)] += 1
    per_stage = {}
    for k, b in acc.items():   # kernels of one stage add up (one launch of each per wave)
        per_stage[stage[k]] = per_stage.get(stage[k], 0.0) + b / cnt[k] / frames
    commit = sys.argv[5] if len(sys.argv) >= 6 else subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    json.dump({"source": "ncu --set full, %s, %d frames per launch" % (sys.argv[1], frames), "commit": commit,
               "dram_bytes_per_frame": per_stage, "total_dram_bytes_per_frame": sum(per_stage.values())}, open(out, "w"), indent=1)
    print(open(out).read())
