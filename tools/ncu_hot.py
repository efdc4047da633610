#!/usr/bin/env python
"""Top sampled SASS instructions of one kernel from an .ncu-rep captured with --import-source on.
usage: python tools/ncu_hot.py gpurun_out/prof.ncu-rep k_seg_build [N]"""
import csv, io, subprocess, sys
rep, kern = sys.argv[1], sys.argv[2]
N = int(sys.argv[3]) if len(sys.argv) > 3 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
body = []
for r in rows[hi + 1:]:
    if r and r[0] == "Kernel Name":
        break          # first captured launch only
    if len(r) == len(hdr):
        body.append(r)
cs, ci, cx = hdr.index("# Samples"), hdr.index("Source"), hdr.index("Instructions Executed")
tot = sum(int(r[cs]) for r in body) or 1
print("kernel %s: %d SASS instructions, %d samples, %d warp-instructions executed" % (kern, len(body), tot, sum(int(r[cx]) for r in body)))
base = int(body[0][0], 16)
top = sorted(range(len(body)), key=lambda i: -int(body[i][cs]))[:N]
for i in sorted(top):
    r = body[i]
    print("%5d  +0x%04x  %5.1f%%  exec %9s  %s" % (i, int(r[0], 16) - base, 100.0 * int(r[cs]) / tot, r[cx], r[ci].strip()))
