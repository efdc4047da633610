#!/usr/bin/env python
"""Measurements beside bench.py's headline line: the other BASELINE.json configs and the widened SURVEY 8(f) rows.
Prints one JSON object per line (not the driver's bench contract).  Run on the B200 box:  python tools/bench_extra.py"""
import json
import os
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _load_pkg import load_pkg, load_synth  # noqa: E402
from bench import tile_batch, FIELDS, algorithmic_bytes  # noqa: E402


def wall(fn, reps):
    fn()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    return (time.perf_counter() - t0) / reps


def main():
    import torch
    pkg, synth = load_pkg(), load_synth()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    # ---- configs[2], [3]: OS1_64 / HDL_32E frames resident in HBM (device path), CUDA events ---------------------------
    for sensor, F, wave in (("OS1_64", 8192, 4096), ("HDL_32E", 16384, 8192)):
        distinct = synth.make_batch(sensor, 32)
        g = pkg.BevGen(sensor, device=0, max_frames_per_batch=wave)
        batch = tile_batch(distinct, F)
        n_total = int(batch["offsets"][-1])
        din = {k: torch.from_numpy(batch[k]).to(dev) for k in FIELDS}
        dout = dict(label=torch.empty((F, g.S), dtype=torch.int16, device=dev), winner=torch.zeros(pkg.winner_words(n_total, F), dtype=torch.int32, device=dev),
                    single=torch.empty((F, 224 * 224), dtype=torch.uint8, device=dev), multi=torch.empty((F, 24 * 224 * 224), dtype=torch.uint8, device=dev))
        pin, pout = {k: v.data_ptr() for k, v in din.items()}, {k: v.data_ptr() for k, v in dout.items()}
        stream = torch.cuda.ExternalStream(g.compute_stream(), device=dev)
        step = lambda: g.process_device(F, batch["offsets"], pin, pout)
        for _ in range(3):
            step()
        g.sync()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for _ in range(5):
            step()
        e1.record(stream); g.sync(); torch.cuda.synchronize()
        ms = e0.elapsed_time(e1) / 5
        alg = algorithmic_bytes(g.S, n_total, F)
        print(json.dumps({"what": "device path, %s" % sensor, "frames_per_s": F / (ms * 1e-3), "us_per_frame": ms * 1e3 / F, "frames_per_step": F,
                          "pts_per_frame": n_total / F, "algorithmic_GBps": alg / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": alg / (ms * 1e-3) / 1e9 / peak}), flush=True)
        g.close(); del din, dout; torch.cuda.empty_cache()
    # ---- 8(f)-1 / 8(f)-3: host-buffer paths through the C-ABI (PCIe inside the timed region), HDL_64E ------------------
    sensor, Fe = "HDL_64E", 256
    distinct = synth.make_batch(sensor, 32)
    hb = tile_batch(distinct, Fe)
    g = pkg.BevGen(sensor, device=0, max_frames_per_batch=64)
    n_total = int(hb["offsets"][-1])
    hin = {}
    for k in FIELDS:
        a = pkg.pinned_empty(hb[k].shape, hb[k].dtype); a[...] = hb[k]; hin[k] = a
    hin["offsets"] = hb["offsets"]
    rec = np.zeros(n_total, np.dtype([("x", "<f4"), ("y", "<f4"), ("z", "<f4"), ("intensity", "<f4"), ("row", "<u2"), ("col", "<u2"), ("t", "<u4"), ("label", "<i2")]))
    for k in FIELDS:
        rec[k] = hb[k]
    prec = pkg.pinned_empty((n_total * 26,), np.uint8); prec[...] = rec.view(np.uint8)
    hout = g.alloc_outputs(Fe, pinned=True, n_total=n_total)
    hout_b = g.alloc_outputs(Fe, pinned=True, n_total=n_total, bvm=True)
    t_soa = wall(lambda: g.process_host(hin, hout), 5)
    t_pack = wall(lambda: g.process_packed_host(prec, hb["offsets"], out=hout), 5)
    t_bvm = wall(lambda: g.process_host(hin, hout_b), 5)
    print(json.dumps({"what": "host path HDL_64E (e2e): SoA staging / packed 26-byte records de-interleaved on the GPU / SoA + bird-view map",
                      "frames_per_s": {"soa": Fe / t_soa, "packed": Fe / t_pack, "soa_with_bvm": Fe / t_bvm},
                      "h2d_bytes_per_frame": {"soa": 22 * n_total / Fe, "packed": 26 * n_total / Fe}}), flush=True)
    # ---- 8(f)-2: projection step, host arrays in/out ---------------------------------------------------------------------
    rng = np.random.default_rng(3)
    n = 8 * 65536
    x = rng.normal(0, 30, n).astype(np.float32); y = rng.normal(0, 30, n).astype(np.float32); z = rng.normal(-1, 3, n).astype(np.float32)
    t_m = wall(lambda: g.project(0, x, y), 5); t_o = wall(lambda: g.project(1, x, y, z), 5)
    print(json.dumps({"what": "bevgen_project, %d points (8 OS1-64 scans), host (pageable) arrays in/out, copies included" % n,
                      "Mpts_per_s": {"mulran": n / t_m / 1e6, "oxford": n / t_o / 1e6}}), flush=True)
    # ---- 8(f)-4: extractTopAndFlatten on one ground-removed HDL_64E cloud (S slots), host arrays in/out -----------------
    S = g.S
    tx = rng.uniform(-100, 100, S).astype(np.float32); ty = rng.uniform(-100, 100, S).astype(np.float32)
    tz = rng.uniform(-2, 10, S).astype(np.float32); tl = (rng.random(S) < 0.5).astype(np.int16)
    t_t = wall(lambda: g.top_flatten(tx, ty, tz, tl), 5)
    print(json.dumps({"what": "bevgen_top_flatten, %d slots (one HDL_64E cloud), host (pageable) arrays in/out, copies included" % S,
                      "clouds_per_s": 1 / t_t, "Mpts_per_s": S / t_t / 1e6}), flush=True)
    # ---- configs[4]: cloud_manip, 2 M points ------------------------------------------------------------------------------
    n = 2_000_000
    hot = rng.random(n) < 0.6
    cx = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    cy = np.where(hot, rng.normal(0, 3, n), rng.uniform(-100, 100, n)).astype(np.float32)
    cz = rng.uniform(-2, 10, n).astype(np.float32)
    th = np.float32(np.deg2rad(37.0)); c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
    t_c = wall(lambda: g.cloud_manip(rt, cx, cy, cz), 5)
    print(json.dumps({"what": "bevgen_cloud_manip, 2 M points (60 % in hot cells), host arrays in/out", "calls_per_s": 1 / t_c, "Mpts_per_s": n / t_c / 1e6}), flush=True)
    g.close()


if __name__ == "__main__":
    main()
