#!/usr/bin/env python
"""A few device-resident bevgen_cloud_manip_device calls on the BASELINE config #5 cloud (2 M points, 60 % of them in a
N(0, 3 m) blob), for an ncu capture of k_cloud_manip:
  ncu --set full --clock-control none --import-source on -k regex:k_cloud_manip -c 3 -f -o gpurun_out/prof_cm python tools/ncu_cloud_manip.py"""
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _load_pkg import load_pkg  # noqa: E402
import torch  # noqa: E402

pkg = load_pkg()
dev = torch.device("cuda", 0)
g = pkg.BevGen("HDL_32E", device=0, max_frames_per_batch=2)
rng = np.random.default_rng(3)
n = 2_000_000
th = np.float32(np.deg2rad(37.0)); c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
blob = rng.random(n) < 0.6
kind = sys.argv[1] if len(sys.argv) > 1 else "blob"
if kind == "uniform":
    cx, cy = rng.uniform(-100, 100, n), rng.uniform(-100, 100, n)
else:
    cx = np.where(blob, rng.normal(0, 3, n), rng.uniform(-100, 100, n)); cy = np.where(blob, rng.normal(0, 3, n), rng.uniform(-100, 100, n))
d = {"x": torch.from_numpy(cx.astype(np.float32)).to(dev), "y": torch.from_numpy(cy.astype(np.float32)).to(dev),
     "z": torch.from_numpy(rng.uniform(-2, 10, n).astype(np.float32)).to(dev)}
for k in ("tx", "ty", "tz"):
    d[k] = torch.empty(n, dtype=torch.float32, device=dev)
d["bev_in"] = torch.empty((201, 201), dtype=torch.float32, device=dev); d["bev_out"] = torch.empty((201, 201), dtype=torch.float32, device=dev)
ptr = {k: v.data_ptr() for k, v in d.items()}
flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
for it in range(4):
    flush.zero_(); torch.cuda.synchronize()
    g.cloud_manip_device(n, rt, ptr); g.sync()
print("done", kind)
g.close()
