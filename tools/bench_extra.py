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


def device_rate(pkg, torch, dev, sensor, distinct, F, wave, device=0, steps=5):
    """frames resident in HBM through bevgen_process_device, CUDA events on the compute stream -> (ms per step, n_total, S)."""
    g = pkg.BevGen(sensor, device=device, max_frames_per_batch=wave, max_points_per_frame=max(int(np.diff(distinct["offsets"]).max()), 1) + 64)
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
    for _ in range(steps):
        step()
    e1.record(stream); g.sync(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    g.set_profiling(True); step(); g.sync(); st = g.stage_ms(); g.set_profiling(False)
    S = g.S
    g.close(); del din, dout; torch.cuda.empty_cache()
    return ms, n_total, S, {k: round(v[0] / F * 1e3, 3) for k, v in st.items() if v[1] > 0}


def sharded():
    """BASELINE configs[2], [3] under torchrun: a 10 000-keyframe OS1_64 / HDL_32E batch sharded by frame index over the
    ranks (one process per GPU, no data-path collective), and the label stage at K = 10 000 with its rows split per rank
    and gathered on the host (here: all_gather of the per-rank row digests)."""
    import hashlib
    import torch
    import torch.distributed as dist
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        sys.stdout.flush(); saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev); dist.barrier(); torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)
    pkg, synth = load_pkg(), load_synth()
    import importlib
    sh = importlib.import_module("pcpt_b200.sharding")     # the shard formulas the gloo test checks on CPU
    K = 10000
    lo, hi = sh.frame_shard(K, rank, world)
    for sensor in ("OS1_64", "HDL_32E"):
        distinct = synth.make_batch(sensor, 32, first=1000 * rank)
        ms, n_total, S, st = device_rate(pkg, torch, dev, sensor, distinct, hi - lo, min(hi - lo, 4096 if sensor == "OS1_64" else 8192), device=local)
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        if rank == 0:
            print(json.dumps({"what": "%s: 10 000 synthetic keyframes sharded over %d GPU(s), device path (32 distinct frames per rank, tiled)" % (sensor, world),
                              "frames_per_s": K / (float(t[0]) * 1e-3), "ms_for_the_batch": float(t[0]), "n_gpus": world, "stage_us_per_frame_rank0": st}), flush=True)
    # label stage: greedy major-frame scan once per rank (serial chain), rows [lo, hi) of the K x M table per rank
    xyz = synth.make_poses(K, seed=11, step=2.0)
    g = pkg.BevGen("OS1_64", device=local, max_frames_per_batch=2)
    t0 = time.perf_counter(); mi, _ = g.select_major(xyz); t_sel = time.perf_counter() - t0
    lo, hi = sh.row_split(K, world)[rank]
    t0 = time.perf_counter(); lab, _, _ = g.labels(xyz, mi, lo, hi); t_lab = time.perf_counter() - t0
    g.close()
    h = hashlib.sha256(np.ascontiguousarray(lab).tobytes()).digest()
    parts = [None] * world
    if world > 1:
        dist.all_gather_object(parts, (lo, hi, lab if K * len(mi) < 8_000_000 else None, h))
    else:
        parts = [(lo, hi, lab, h)]
    if rank == 0:
        meta = json.load(open(os.path.join(ROOT, "tests", "golden", "bev_golden.json")))["labels"]["K10000_s11"]
        full = np.concatenate([p[2] for p in parts]) if all(p[2] is not None for p in parts) else None
        ok = None
        if full is not None:
            sys.path.insert(0, os.path.join(ROOT, "tests")); import cases
            ok = cases.digest(full) == meta["labels"] and len(mi) == meta["M"]
        print(json.dumps({"what": "label stage K = 10 000, M = %d, rows split over %d rank(s)" % (len(mi), world), "select_major_s": t_sel, "labels_rows_s_rank0": t_lab,
                          "equals_reference_generated_golden": ok}), flush=True)
    if world > 1:
        dist.destroy_process_group()


def contention(pkg, torch, dev, synth, peak):
    """does a heavily contended cell need a sort-by-cell fallback?  (a) every point of a full HDL_64E frame in
    ONE BEV cell vs the synthetic frames, device path; (b) cloud_manip on 2 M device-resident points: uniform, the
    config #5 blob (60 % of the points within ~3 m), and all points in one cell."""
    sys.path.insert(0, os.path.join(ROOT, "tests")); import cases
    sensor = "HDL_64E"
    sp = pkg.sensor_params(sensor)
    class SP: n_scan, horizon_scan, S = sp.n_scan, sp.horizon_scan, sp.S
    hot = cat = None
    frames = [cases.hot_cell_frame(SP, seed=s) for s in range(8)]
    offs = np.zeros(9, np.int64); offs[1:] = np.cumsum([len(f["x"]) for f in frames])
    hot = {k: np.concatenate([f[k] for f in frames]) for k in FIELDS}; hot["offsets"] = offs
    uni = synth.make_batch(sensor, 8)
    res = {}
    for name, d in (("synthetic", uni), ("one_cell", hot)):
        ms, n_total, S, st = device_rate(pkg, torch, dev, sensor, d, 2220, 2220)
        res[name] = {"us_per_frame": ms * 1e3 / 2220, "pts_per_frame": n_total / 2220, "stage_us_per_frame": st}
    res["finalize_ratio_one_cell_vs_synthetic"] = res["one_cell"]["stage_us_per_frame"]["finalize_bin_scatter"] / res["synthetic"]["stage_us_per_frame"]["finalize_bin_scatter"]
    print(json.dumps({"what": "contention: all %d points of a frame in one BEV cell vs synthetic frames (device path, HDL_64E)" % SP.S, **res}), flush=True)
    g = pkg.BevGen("HDL_32E", device=0, max_frames_per_batch=2)
    rng = np.random.default_rng(3)
    n = 2_000_000
    th = np.float32(np.deg2rad(37.0)); c, s = np.float32(np.cos(th)), np.float32(np.sin(th))
    rt = np.array([c, -s, 0, 3.5, s, c, 0, -1.25, 0, 0, 1, 0.2], np.float32)
    blob = rng.random(n) < 0.6
    clouds = {"uniform": (rng.uniform(-100, 100, n), rng.uniform(-100, 100, n)),
              "config5_blob": (np.where(blob, rng.normal(0, 3, n), rng.uniform(-100, 100, n)), np.where(blob, rng.normal(0, 3, n), rng.uniform(-100, 100, n))),
              "one_cell": (rng.uniform(10.1, 10.9, n), rng.uniform(-7.9, -7.1, n))}
    stream = torch.cuda.ExternalStream(g.compute_stream(), device=dev)
    out = {}
    for name, (cx, cy) in clouds.items():
        d = {"x": torch.from_numpy(cx.astype(np.float32)).to(dev), "y": torch.from_numpy(cy.astype(np.float32)).to(dev),
             "z": torch.from_numpy(rng.uniform(-2, 10, n).astype(np.float32)).to(dev)}
        for k in ("tx", "ty", "tz"):
            d[k] = torch.empty(n, dtype=torch.float32, device=dev)
        d["bev_in"] = torch.empty((201, 201), dtype=torch.float32, device=dev); d["bev_out"] = torch.empty((201, 201), dtype=torch.float32, device=dev)
        ptr = {k: v.data_ptr() for k, v in d.items()}
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        ts = []
        for it in range(8):
            flush.zero_()                               # 24 MB of points fit the L2: flush it between iterations
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(stream); g.cloud_manip_device(n, rt, ptr); e1.record(stream); g.sync(); torch.cuda.synchronize()
            if it >= 3:
                ts.append(e0.elapsed_time(e1))
        ms = float(np.median(ts))
        out[name] = {"us_per_call": ms * 1e3, "calls_per_s": 1e3 / ms, "algorithmic_GBps": 48323208 / (ms * 1e-3) / 1e9, "frac_of_measured_hbm_peak": 48323208 / (ms * 1e-3) / 1e9 / peak}
        # back to back: 24 calls enqueued without a host sync, rotating over 4 copies of the cloud and of the outputs (192 MB per
        # round > L2), so the launch gaps of a single call (memset + kernel + merge issued from Python) drop out
        sets = [ptr]
        keep = [d]
        for _ in range(3):
            e = {k: v.clone() for k, v in d.items()}
            keep.append(e); sets.append({k: v.data_ptr() for k, v in e.items()})
        for q in range(4):
            g.cloud_manip_device(n, rt, sets[q])
        g.sync(); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for q in range(24):
            g.cloud_manip_device(n, rt, sets[q % 4])
        e1.record(stream); g.sync(); torch.cuda.synchronize()
        msb = e0.elapsed_time(e1) / 24
        out[name]["back_to_back_us_per_call"] = msb * 1e3
        out[name]["back_to_back_frac_of_measured_hbm_peak"] = 48323208 / (msb * 1e-3) / 1e9 / peak
        del d, flush, keep
    out["ratio_one_cell_vs_uniform"] = out["one_cell"]["us_per_call"] / out["uniform"]["us_per_call"]
    out["ratio_blob_vs_uniform"] = out["config5_blob"]["us_per_call"] / out["uniform"]["us_per_call"]
    print(json.dumps({"what": "cloud_manip (config #5), 2 M points resident in HBM, L2 flushed between calls, 48 323 208 algorithmic bytes per call (incl. the two memsets + kernel)", **out}), flush=True)
    g.close()


def main():
    import torch
    if "--sharded" in sys.argv:
        return sharded()
    pkg, synth = load_pkg(), load_synth()
    dev = torch.device("cuda", 0)
    peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))).get("hbm_gbs", 6650.0) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else 6650.0
    if "--contention" in sys.argv:
        return contention(pkg, torch, dev, synth, peak)
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
    contention(pkg, torch, dev, synth, peak)
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
    # ---- the single-frame form: bevgen_submit / bevgen_collect with 8 frames in flight ---------------
    g8 = pkg.BevGen(sensor, device=0, max_frames_per_batch=8)
    fl = [{k: hb[k][int(hb["offsets"][i]):int(hb["offsets"][i + 1])] for k in FIELDS} for i in range(32)]
    nfr = 256
    for i in range(8):
        g8.submit(i, fl[i % 32])
    t0 = time.perf_counter()
    for i in range(nfr):
        g8.collect(i)
        if i + 8 < nfr + 8:
            g8.submit(i + 8, fl[(i + 8) % 32])
    dt = time.perf_counter() - t0
    for i in range(nfr, nfr + 8):
        g8.collect(i)
    g8.close()
    print(json.dumps({"what": "bevgen_submit / bevgen_collect, one HDL_64E frame per call, 8 in flight, pageable host arrays, Python caller (collect allocates its outputs)",
                      "frames_per_s": nfr / dt}), flush=True)
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
