#!/usr/bin/env python
"""bench.py — BEV frames/s of the batch_multi_bev_gen hot path on B200 (BASELINE.json metric).

  python bench.py --gpus N --steps K --warmup W            our CUDA path (C-ABI, libbevgen_cuda.so)
  python bench.py --impl reference ...                     the reference algorithm on the host cores (CPU oracle port)

A "step" = one pass of the hot path (getOrderedCloud + markGroundPoints + single & multi BEV,
BatchMultiBevGen.cpp:735-747, no file encoders) over one batch of synthetic HDL_64E keyframes (BASELINE configs[1]).
  value : frames/s with the batch already resident in HBM (bevgen_process_device), CUDA events on the compute stream
  e2e   : frames/s through bevgen_process_host_compact (the C-ABI call a host makes: 16 B/point staging format in, ground
          bits + bit planes out) with pinned HOST buffers, H2D and D2H inside the timed region; e2e.full_layout is the
          same through bevgen_process_host (22 B/point SoA in, reference-layout bytes out); e2e.pcie_alone is the copy
          engines moving the same bytes with no kernels, which names the limiter
  roofline : dominant kernel, algorithmic bytes per launch / its mean launch duration (CUDA events in the library)
  parity_checked : outputs of the TIMED runs compared with the oracle outside the timed region (frames at the head, at a
          wave boundary and at the tail of the device batch; frames of the e2e batch after host-side expansion)
  cpu_baseline : the oracle port on all host threads + the reference's own source (oracle/_ref, one process per core),
          bounded samples, rank 0 / N=1 only; cli: the drop-in CLI and the reference's main() on a keyframe folder
Multi-GPU: frames shard by index, one process per GPU, no data-path collective ("weak" scaling: F frames per rank);
torch.distributed is only used for the barrier and the max-over-ranks of the timed region.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
from _load_pkg import load_pkg, load_synth, load_oracle  # noqa: E402

FIELDS = ("x", "y", "z", "intensity", "row", "col", "label")
METRIC = "BEV frames/sec (HDL-64E)"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--sensor", default="HDL_64E")
    ap.add_argument("--frames", type=int, default=4440, help="frames resident per GPU per step (device path)")
    ap.add_argument("--e2e-frames", type=int, default=1024, help="frames per step of the host-buffer (e2e) path")
    ap.add_argument("--distinct", type=int, default=32, help="distinct synthetic frames generated, then tiled")
    ap.add_argument("--wave", type=int, default=int(os.environ.get("BEVGEN_WAVE", "2220")), help="frames per launch wave")
    ap.add_argument("--ref-frames", type=int, default=0, help="reference arm: frames per step (0 = 32 per host thread)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-cli", action="store_true", help="skip the CLI / reference main() folder runs (rank 0, N=1)")
    ap.add_argument("--no-parity", action="store_true")
    ap.add_argument("--wc", action="store_true", help="e2e input staging buffers write-combined (bevgen_host_alloc_wc)")
    return ap.parse_args()


def tile_batch(distinct, F):
    """Tile the distinct frames to F frames (concatenated SoA + offsets)."""
    offs_d = distinct["offsets"]
    D = len(offs_d) - 1
    order = [i % D for i in range(F)]
    lens = np.array([offs_d[i + 1] - offs_d[i] for i in order], np.int64)
    offs = np.zeros(F + 1, np.int64); offs[1:] = np.cumsum(lens)
    out = {}
    for k in FIELDS:
        out[k] = np.concatenate([distinct[k][offs_d[i]:offs_d[i + 1]] for i in order])
    out["offsets"] = offs
    return out


class ClockSampler(threading.Thread):
    """SM clock + throttle reasons DURING the timed region (B200_PROFILING.md recipe).  NVML (nvidia_ml_py) answers in
    well under a millisecond, so a 75 ms timed region still yields a dozen samples; `nvidia-smi` (50-100 ms per call) is the
    fallback when the module or the query is unavailable."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, dev):
        super().__init__(daemon=True)
        self.dev, self.rows, self.stop_ev = dev, [], threading.Event()   # rows: (sm_mhz, max_mhz, {reasons})
        self.mem, self.power = [], []
        self.nv = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            idx = int(vis.split(",")[dev]) if vis and all(t.strip().isdigit() for t in vis.split(",")) else dev
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_mhz = float(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
            self.nv = pynvml
        except Exception:
            self.nv = None

    def _nvml_sample(self):
        nv = self.nv
        sm = float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM))
        try:
            self.mem.append(float(nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_MEM)))
            self.power.append(nv.nvmlDeviceGetPowerUsage(self.h) / 1e3)
        except Exception:
            pass
        get = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
        mask = int(get(self.h))
        names = {"hw_slowdown": 0x8, "sw_thermal_slowdown": 0x20, "hw_thermal_slowdown": 0x40, "sw_power_cap": 0x4}
        return sm, self.max_mhz, {n for n, b in names.items() if mask & b}

    def _smi_sample(self):
        o = subprocess.run(["nvidia-smi", "-i", str(self.dev), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                           capture_output=True, text=True, timeout=5).stdout.strip()
        r = [c.strip() for c in o.split(",")]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        return float(r[1]), float(r[2]), {n for n, v in zip(names, r[5:9]) if v.lower().startswith("active")}

    def run(self):
        while not self.stop_ev.is_set():
            try:
                self.rows.append(self._nvml_sample() if self.nv else self._smi_sample())
            except Exception:
                if self.nv:
                    self.nv = None      # fall back to nvidia-smi
                    continue
            self.stop_ev.wait(0.004 if self.nv else 0.2)

    def summary(self):
        self.stop_ev.set(); self.join(timeout=6)
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock query unavailable"]}
        sm = sorted(r[0] for r in self.rows)
        reasons = sorted(set().union(*[r[2] for r in self.rows]))
        out = {"sm_mhz": sm[len(sm) // 2], "sm_max_mhz": self.rows[0][1], "reasons": reasons, "samples": len(self.rows),
               "source": "nvml" if self.nv else "nvidia-smi"}
        if self.mem:
            out["mem_mhz"] = sorted(self.mem)[len(self.mem) // 2]
        if self.power:
            out["power_w"] = sorted(self.power)[len(self.power) // 2]
        return out


def algorithmic_bytes(sensor_S, n_in_total, F):
    # n_in*22 [x,y,z,intensity f32 + row,col u16 + label i16] + S*2 [labels] + 50176 [single] + 1204224 [multi]
    return n_in_total * 22 + F * (sensor_S * 2 + 50176 + 1204224)


def workload_config(args, world, n_total, in_bytes):
    """`config` of the bench line, the same dict in both arms (the reference arm times bounded samples of this workload: its
    sample size is in `cpu_baseline.sample` / `sample_frames_per_step`).  n_total = points of rank 0's F tiled frames."""
    F = args.frames
    return {"workload": "BASELINE configs[1]: %s synthetic keyframes (~%dk pts/frame), hot loop BatchMultiBevGen.cpp:735-747"
                        % (args.sensor, round(n_total / F / 1000)),
            "sensor": args.sensor, "frames_per_step_per_gpu": F, "frames_per_wave": args.wave,
            "l2_policy": "inputs larger than L2 (%.2f GB of points per step)" % (in_bytes / 1e9),
            "parallelism": "frames sharded by index, %d process(es), no collective" % world}


def ref_source_rate(O, args, distinct, cores, seconds):
    """frames/s of the reference's OWN source (oracle/_ref/libbevgen_ref.so = BatchMultiBevGen.cpp compiled against
    oracle/stub), one process per host core; None when the library did not travel to this box."""
    if O.ref_bevgen_lib() is None:
        return None
    d = "/dev/shm" if os.access("/dev/shm", os.W_OK) else "/tmp"
    path = os.path.join(d, "bevgen_bench_%d.npz" % os.getpid())
    sub = {k: distinct[k] for k in FIELDS}; sub["offsets"] = distinct["offsets"]
    np.savez(path, **sub)
    try:
        return O.ref_bench_all_cores(args.sensor, path, cores, seconds=seconds)
    finally:
        os.remove(path)


def run_reference(args, rank, world):
    """Reference arm: the reference's CPU implementation of the hot loop body on all host cores, bounded sample of the same
    workload.  Two measurements: `reference` = the reference's own source text (oracle/_ref, built from
    /root/reference/BatchMultiBevGen.cpp against stand-in PCL/OpenCV headers; PNG / CSV encoders stubbed out, one process
    per core because of its file-scope globals) and `port` = the oracle restatement in C (threads).  The line's value is
    the reference-source one when that library is present, else the port.  Rank 0 only."""
    if rank != 0:
        return
    O, synth = load_oracle(), load_synth()
    cores = os.cpu_count() or 1
    F = args.ref_frames or 32 * cores
    full = synth.make_batch(args.sensor, args.distinct)                         # rank 0's distinct frames of the product arm
    distinct = full if F >= args.distinct else synth.make_batch(args.sensor, F)
    lens = np.diff(full["offsets"])
    n_tiled = int(sum(int(lens[i % len(lens)]) for i in range(args.frames)))   # points of the product arm's args.frames tiled frames (22 B each)
    batch = tile_batch(distinct, F)
    sp = O.sensor(args.sensor)
    call = lambda: O.frames(sp, batch["offsets"], *[batch[k] for k in FIELDS], n_threads=cores)
    for _ in range(min(args.warmup, 1)):
        call()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        call()
    dt = time.perf_counter() - t0
    v_port = F * args.steps / dt
    rs = ref_source_rate(O, args, distinct, cores, seconds=min(max(dt, 6.0), 20.0))
    v = rs["frames_per_s"] if rs else v_port
    kind = "reference" if rs else "port"
    sample = ("%d processes x ref_bench over %d distinct frames for %.1f s (%d frames); reference source, encoders stubbed" %
              (rs["procs"], len(distinct["offsets"]) - 1, rs["wall_s"], rs["frames"])) if rs else \
             ("%d frames/step x %d steps, %d threads (one frame per thread at a time)" % (F, args.steps, cores))
    line = {"impl": "reference", "metric": METRIC, "value": v, "unit": "frames/s", "n_gpus": args.gpus, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": (F / v) * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world, n_tiled, 22 * n_tiled), "sample_frames_per_step": F,
            "cpu_baseline": {"value": v, "unit": "frames/s", "cores": cores, "kind": kind, "sample": sample,
                             "port_frames_per_s": v_port, "reference_source": rs},
            "e2e": {"value": v, "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def check_device_outputs(pkg, g, O, args, distinct, offs, dout, F):
    """parity of the TIMED device-resident run: frames at the head, around the first wave boundary and at the tail of the
    batch are pulled back and compared with the oracle (the batch is `distinct` frames tiled, frame i = distinct i % D)."""
    D = len(distinct["offsets"]) - 1
    sp = O.sensor(args.sensor)
    ref = O.frames(sp, distinct["offsets"], *[distinct[k] for k in FIELDS], n_threads=os.cpu_count() or 1)
    W = min(args.wave, F)
    idx = sorted(set(list(range(min(D, F))) + [i for i in range(W - 8, W + 8) if 0 <= i < F] + list(range(max(F - D, 0), F))))
    winner = dout["winner"].cpu().numpy().view(np.uint32)
    bad = []
    for i in idx:
        d = i % D
        n = int(distinct["offsets"][d + 1] - distinct["offsets"][d])
        got = dict(label=dout["label"][i].cpu().numpy(), single=dout["single"][i].cpu().numpy().reshape(224, 224),
                   multi=dout["multi"][i].cpu().numpy().reshape(24, 224, 224))
        ok = all(np.array_equal(got[k], ref[k][d]) for k in got)
        want = np.zeros(n, bool); want[ref["owner"][d][ref["owner"][d] > 0].astype(np.int64) - 1] = True   # last writers of their slots
        ok = ok and np.array_equal(pkg.winner_mask(winner, offs, i), want)
        if not ok:
            bad.append(i)
    return {"frames": len(idx), "ok": not bad, "bad_frames": bad[:8], "against": "oracle (pinned to the reference source, tests/test_reference_source_pin.py)"}


def cli_numbers(pkg, synth, O, args, n_cli=100, n_ref=12):
    """End-to-end numbers a user of the drop-in sees (rank 0, N=1): the CLI on an n_cli-keyframe folder with every file
    written and with --no-encode, and the reference's own main() (oracle/_ref) on n_ref keyframes - its
    "[TIME] Average preprocessing and BEV generation" print is BASELINE.md's C1 (order -> BEV files, 1 core)."""
    import importlib, shutil, tempfile
    pcd = importlib.import_module("pcpt_b200.pcd")
    base = tempfile.mkdtemp(prefix="bevgen_cli_", dir="/dev/shm" if os.access("/dev/shm", os.W_OK) else None)
    out = {}
    try:
        root = os.path.join(base, "kf"); os.makedirs(os.path.join(root, "keyframe_point_cloud"))
        fr = [synth.make_frame(args.sensor, 5000 + i) for i in range(min(n_cli, 32))]
        for i in range(n_cli):
            pcd.write(os.path.join(root, "keyframe_point_cloud", "%06d.pcd" % i), fr[i % len(fr)])
        open(os.path.join(root, "keyframe_pose.csv"), "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n_cli, seed=5, step=9.0))) + "\n")
        for tag, extra in (("all_files", []), ("no_encode", ["--no-encode", "--no-pcd"])):
            mj = os.path.join(base, tag + ".json")
            t0 = time.perf_counter()
            r = subprocess.run([pkg.CLI_PATH, root, args.sensor, "--json-metrics", mj] + extra, capture_output=True, text=True, timeout=900)
            wall = time.perf_counter() - t0
            if r.returncode != 0:
                out[tag] = {"error": r.stderr[-300:]}
                continue
            m = json.load(open(mj)) if os.path.exists(mj) else {}
            out[tag] = {"frames": n_cli, "process_wall_s": wall, "frames_per_s_process": n_cli / wall, "metrics": m}
        # the same with the folder grown to n_long keyframes: context creation and the pinned staging sets (tens of milliseconds each)
        # are most of a 100-keyframe run; this is the rate a long folder sees
        n_long = 600
        try:
            for i in range(n_cli, n_long):
                pcd.write(os.path.join(root, "keyframe_point_cloud", "%06d.pcd" % i), fr[i % len(fr)])
            open(os.path.join(root, "keyframe_pose.csv"), "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n_long, seed=5, step=9.0))) + "\n")
            mj = os.path.join(base, "long.json")
            t0 = time.perf_counter()
            r = subprocess.run([pkg.CLI_PATH, root, args.sensor, "--json-metrics", mj], capture_output=True, text=True, timeout=900)
            wall = time.perf_counter() - t0
            out["all_files_%d_keyframes" % n_long] = ({"error": r.stderr[-300:]} if r.returncode != 0 else
                                                      {"frames": n_long, "process_wall_s": wall, "frames_per_s_process": n_long / wall,
                                                       "metrics": json.load(open(mj)) if os.path.exists(mj) else {}})
        except Exception as e:    # e.g. a small /dev/shm: this leg must never take the bench line down
            out["all_files_%d_keyframes" % n_long] = {"error": repr(e)}
        if O.ref_bevgen_lib() is not None:
            rroot = os.path.join(base, "ref"); os.makedirs(os.path.join(rroot, "keyframe_point_cloud"))
            for i in range(n_ref):
                shutil.copy(os.path.join(root, "keyframe_point_cloud", "%06d.pcd" % i), os.path.join(rroot, "keyframe_point_cloud"))
            open(os.path.join(rroot, "keyframe_pose.csv"), "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n_ref, seed=5, step=9.0))) + "\n")
            t0 = time.perf_counter()
            rc, so, se = O.ref_main(rroot, args.sensor)
            wall = time.perf_counter() - t0
            avg = [l for l in so.splitlines() if l.startswith("[TIME] Average preprocessing and BEV generation:")]
            out["reference_main"] = {"frames": n_ref, "rc": rc, "process_wall_s": wall, "frames_per_s_process": n_ref / wall,
                                     "C1_ms_per_frame_timed_span": float(avg[0].split(":")[1]) if avg else None, "cores": 1,
                                     "note": "the reference's own main() (BatchMultiBevGen.cpp:664-771) built against oracle/stub; span = order + ground + both BEVs incl. "
                                             ".bin / 25 PNG / CSV writes and the per-frame mkdir fork, like its [TIME] print"}
    finally:
        shutil.rmtree(base, ignore_errors=True)
    return out


def main():
    args = parse()
    rank = int(os.environ.get("RANK", "0")); world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank, world)

    import torch
    import torch.distributed as dist
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the product path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        # NCCL prints its version banner to stdout when the first communicator is built; stdout must carry exactly one
        # JSON line, so fd 1 points at stderr until the communicator exists
        sys.stdout.flush()
        saved = os.dup(1); os.dup2(2, 1)
        try:
            dist.init_process_group("nccl", device_id=dev)
            dist.barrier()
            torch.cuda.synchronize()
        finally:
            sys.stdout.flush(); os.dup2(saved, 1); os.close(saved)

    pkg, synth = load_pkg(), load_synth()
    F, Fe = args.frames, args.e2e_frames
    distinct = synth.make_batch(args.sensor, args.distinct, first=1000 * rank)   # each rank its own shard of keyframes
    g = pkg.BevGen(args.sensor, device=local, max_frames_per_batch=args.wave)
    S = g.S

    # ---- device-resident batch (value) ------------------------------------------------------------------------
    batch = tile_batch(distinct, F)
    n_total = int(batch["offsets"][-1])
    din = {k: torch.from_numpy(batch[k]).to(dev) for k in FIELDS}
    dout = dict(label=torch.empty((F, S), dtype=torch.int16, device=dev),
                winner=torch.zeros(pkg.winner_words(n_total, F), dtype=torch.int32, device=dev),
                single=torch.empty((F, 224 * 224), dtype=torch.uint8, device=dev),
                multi=torch.empty((F, 24 * 224 * 224), dtype=torch.uint8, device=dev))
    pin, pout = {k: v.data_ptr() for k, v in din.items()}, {k: v.data_ptr() for k, v in dout.items()}
    in_bytes = sum(v.numel() * v.element_size() for v in din.values())
    stream = torch.cuda.ExternalStream(g.compute_stream(), device=dev)
    offs = batch["offsets"]
    del batch                      # the points now live in HBM; keep the host footprint small (8 ranks share one box)
    step = lambda: g.process_device(F, offs, pin, pout)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, steps, on_stream):
        """EXACTLY `steps` calls bracketed by barrier + synchronize; CUDA events on the launching stream.
        Returns (event ms, wall ms) as the max over ranks, then this rank's own pair."""
        barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t0 = time.perf_counter()
        e0.record(on_stream)
        for _ in range(steps):
            fn()
        e1.record(on_stream)
        g.sync(); torch.cuda.synchronize()
        wall = (time.perf_counter() - t0) * 1e3
        ms = e0.elapsed_time(e1)
        barrier()
        t = torch.tensor([ms, wall], device=dev, dtype=torch.float64)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t[0]), float(t[1]), ms, wall

    for _ in range(max(args.warmup, 3)):
        step()
    g.sync()
    barrier()      # the first NCCL barrier builds the communicator (~1 s): keep it out of the sampled / timed region
    sampler = ClockSampler(local); sampler.start()
    l0 = g.kernel_launches()
    ms, wall_ms, ms_local, _ = timed(step, args.steps, stream)
    launches = g.kernel_launches() - l0
    clocks = sampler.summary()
    value = world * F * args.steps / (ms * 1e-3)

    parity = None
    if not args.no_parity:     # outputs of the run just timed, against the oracle, outside the timed region
        parity = {"device_path": check_device_outputs(pkg, g, load_oracle(), args, distinct, offs, dout, F)}

    # ---- per-stage CUDA-event timing of the same step (roofline of the dominant kernel) --------------------------
    g.set_profiling(True)
    for _ in range(args.steps):
        step()
    g.sync()
    st = g.stage_ms()
    g.set_profiling(False)
    kern = {k: v for k, v in st.items() if v[1] > 0 and k != "clear"}
    dom = max(kern, key=lambda k: kern[k][0])
    dom_ms_per_launch = kern[dom][0] / kern[dom][1]
    frames_per_launch = F * args.steps / kern[dom][1]
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = float(peaks.get("hbm_gbs", 6650.0)); peak_src = "measured (MEASURED_PEAKS.json hbm_gbs)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
    alg = algorithmic_bytes(S, n_total, F) / F * frames_per_launch
    achieved = alg / (dom_ms_per_launch * 1e-3) / 1e9
    # measured DRAM bytes of the dominant kernel per launch: one `ncu --set full` capture of this command at a smaller
    # wave (tools/ncu_summary.py -> profiles/dominant_kernel_traffic.json: bytes per frame + the commit it was taken at)
    traffic, traffic_src = None, None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "dominant_kernel_traffic.json")))
        if args.sensor == "HDL_64E" and dom in tj["dram_bytes_per_frame"]:
            traffic = tj["dram_bytes_per_frame"][dom] * frames_per_launch
            traffic_src = {"file": "profiles/dominant_kernel_traffic.json", "capture": tj.get("source"), "commit": tj.get("commit"),
                           "all_kernels_dram_bytes_per_frame": tj.get("total_dram_bytes_per_frame")}
    except Exception:
        pass
    whole = algorithmic_bytes(S, n_total, F) * args.steps / (ms * 1e-3) / 1e9
    roofline = {"bound": "hbm", "kernel": dom, "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": traffic, "traffic_source": traffic_src, "peak_source": peak_src,
                "algorithmic_bytes_per_frame": alg / frames_per_launch,
                "frames_per_launch": frames_per_launch, "ms_per_launch": dom_ms_per_launch,
                "whole_path_GBps": whole, "whole_path_frac": whole / peak,
                "stage_ms_per_step": {k: v[0] / args.steps for k, v in st.items() if v[1] > 0},
                "stage_us_per_frame": {k: v[0] / args.steps / F * 1e3 for k, v in st.items() if v[1] > 0}}

    # ---- e2e: host buffers through the C-ABI, pinned, H2D + D2H inside the timed region ------------------------------
    del din, dout
    torch.cuda.empty_cache()
    hb = tile_batch(distinct, Fe)
    e_offs = hb["offsets"]

    def pin_copy(a):      # input staging: written once by the host, then only read by the copy engine
        p = pkg.pinned_empty(a.shape, a.dtype, write_combined=args.wc); p[...] = a
        return p
    # (a) compact staging format (bevgen_process_host_compact): x, y, z + (slot | flags) in; ground bits, winner bits,
    #     single and the 3 occupancy bit planes out.  The meta word is what a PCD parser writes instead of four SoA fields.
    cin = {k: pin_copy(hb[k]) for k in ("x", "y", "z")}
    cin["meta"] = pin_copy(pkg.pack_meta(g.params, hb["row"], hb["col"], hb["intensity"], hb["label"]))
    cin["offsets"] = e_offs
    cout = g.alloc_outputs_compact(Fe, pinned=True, n_total=int(e_offs[-1]))
    h2d = int(sum(cin[k].nbytes for k in ("x", "y", "z", "meta"))) + e_offs.nbytes
    d2h = int(sum(v.nbytes for v in cout.values()))
    cstep = lambda: g.process_host_compact(cin, cout)
    for _ in range(2):
        cstep()
    # process_host* return only when the outputs are in host memory, so the wall clock brackets the device work;
    # events on torch's current stream would not see the library's three streams.
    _, e_wall, _, e_wall_local = timed(cstep, args.steps, torch.cuda.current_stream())
    e2e_v = world * Fe * args.steps / (e_wall * 1e-3)
    if parity is not None:   # the bytes the timed e2e run left in host memory, expanded on the host, against the oracle
        O = load_oracle()
        D = len(distinct["offsets"]) - 1
        ref = O.frames(O.sensor(args.sensor), distinct["offsets"], *[distinct[k] for k in FIELDS], n_threads=os.cpu_count() or 1)
        pick = sorted(set(list(range(min(Fe, 96))) + list(range(max(Fe - 96, 0), Fe)) + list(range(0, Fe, max(Fe // 64, 1)))))   # head, tail, every 16th
        exp = g.compact_to_reference_layout(cout, hb, frames=pick)
        bad = [i for j, i in enumerate(pick) if not all(np.array_equal(exp[k][j], ref[k][i % D]) for k in ("owner", "label", "single", "multi"))]
        parity["e2e_compact"] = {"frames": len(pick), "of": Fe, "ok": not bad, "bad_frames": bad[:8]}
        del exp
    # (b) the copy engines alone on the same bytes, both directions at once, no kernels: what PCIe gives this process
    dbuf = torch.empty(max(h2d, d2h) + 64, dtype=torch.uint8, device=dev)
    hsrc = torch.empty(h2d, dtype=torch.uint8).pin_memory(); hdst = torch.empty(d2h, dtype=torch.uint8).pin_memory()
    s1, s2 = torch.cuda.Stream(device=dev), torch.cuda.Stream(device=dev)

    def copies():
        with torch.cuda.stream(s1):
            dbuf[:h2d].copy_(hsrc, non_blocking=True)
        with torch.cuda.stream(s2):
            hdst.copy_(dbuf[:d2h], non_blocking=True)
        s1.synchronize(); s2.synchronize()
    copies()
    _, p_wall, _, p_wall_local = timed(copies, args.steps, torch.cuda.current_stream())
    del dbuf, hsrc, hdst
    for a in [cin[k] for k in ("x", "y", "z", "meta")] + list(cout.values()):
        pkg.pinned_free(a)
    # (c) the reference-layout entry point (bevgen_process_host: 22 B/point SoA in, labels i16 + 24 expanded layers out)
    hin = {k: pin_copy(hb[k]) for k in FIELDS}
    hin["offsets"] = e_offs
    hout = g.alloc_outputs(Fe, pinned=True, n_total=int(e_offs[-1]))
    h2d_full = int(sum(hin[k].nbytes for k in FIELDS)) + e_offs.nbytes
    d2h_full = int(sum(v.nbytes for v in hout.values()))
    estep = lambda: g.process_host(hin, hout)
    for _ in range(2):
        estep()
    _, f_wall, _, _ = timed(estep, args.steps, torch.cuda.current_stream())
    e2e_full_v = world * Fe * args.steps / (f_wall * 1e-3)
    for a in [hin[k] for k in FIELDS] + list(hout.values()):
        pkg.pinned_free(a)
    del hb

    # ---- per-rank record: who is slow, and on what -------------------------------------------------------------------
    stage_names = [k for k, v in st.items() if v[1] > 0]
    mine = torch.tensor([ms_local / args.steps, e_wall_local / args.steps, h2d * args.steps / (e_wall_local * 1e-3) / 1e9,
                         h2d * args.steps / (p_wall_local * 1e-3) / 1e9, d2h * args.steps / (p_wall_local * 1e-3) / 1e9,
                         float(clocks.get("sm_mhz") or 0), float(clocks.get("mem_mhz") or 0), float(clocks.get("power_w") or 0)] +
                        [st[k][0] / args.steps / F * 1e3 for k in stage_names], device=dev, dtype=torch.float64)
    allr = [torch.zeros_like(mine) for _ in range(world)]
    if world > 1:
        dist.all_gather(allr, mine)
    else:
        allr = [mine]
    per_rank = [{"rank": r, "ms_per_step": float(t[0]), "e2e_ms_per_step": float(t[1]), "e2e_h2d_GBps": float(t[2]),
                 "pcie_alone_h2d_GBps": float(t[3]), "pcie_alone_d2h_GBps": float(t[4]), "sm_mhz": float(t[5]), "mem_mhz": float(t[6]),
                 "power_w": float(t[7]), "stage_us_per_frame": {k: round(float(t[8 + i]), 4) for i, k in enumerate(stage_names)}}
                for r, t in enumerate(allr)]
    slow = max(per_rank, key=lambda r: r["ms_per_step"]); fast = min(per_rank, key=lambda r: r["ms_per_step"])

    # ---- CPU baseline (rank 0, N=1 only) --------------------------------------------------------------------------
    cpu, cli = None, None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        O = load_oracle()
        cores = os.cpu_count() or 1
        sp = O.sensor(args.sensor)
        nfr = min(16 * cores, 1024)
        cb = tile_batch(distinct, nfr)
        reps, t0 = 0, time.perf_counter()
        while True:   # bounded sample: ~8 s of wall time on all host threads
            O.frames(sp, cb["offsets"], *[cb[k] for k in FIELDS], n_threads=cores)
            reps += 1
            dt = time.perf_counter() - t0
            if dt >= 8.0 or reps >= 200:
                break
        t1 = time.perf_counter()
        one = tile_batch(distinct, min(16, args.distinct))
        O.frames(sp, one["offsets"], *[one[k] for k in FIELDS], n_threads=1)
        dt1 = time.perf_counter() - t1
        rs = ref_source_rate(O, args, distinct, cores, seconds=8.0)
        v_port = nfr * reps / dt
        cpu = {"value": rs["frames_per_s"] if rs else v_port, "unit": "frames/s", "cores": cores, "kind": "reference" if rs else "port",
               "sample": ("reference source (oracle/_ref, encoders stubbed): %d processes x %d distinct frames for %.1f s; " % (rs["procs"], args.distinct, rs["wall_s"]) if rs else "") +
                         "port: %d frames x %d passes on %d threads (%.1f s); single-thread port: %.1f frames/s over %d frames" %
                         (nfr, reps, cores, dt, (len(one["offsets"]) - 1) / dt1, len(one["offsets"]) - 1),
               "port_frames_per_s": v_port, "port_single_thread_frames_per_s": (len(one["offsets"]) - 1) / dt1, "reference_source": rs}
        try:    # cloud_manip on 2 M points (config #5's cloud), the oracle port on one core
            rng = np.random.default_rng(3); n_cm = 2_000_000
            blob = rng.random(n_cm) < 0.6
            cx = np.where(blob, rng.normal(0, 3, n_cm), rng.uniform(-100, 100, n_cm)).astype(np.float32)
            cy = np.where(blob, rng.normal(0, 3, n_cm), rng.uniform(-100, 100, n_cm)).astype(np.float32)
            cz = rng.uniform(-2, 10, n_cm).astype(np.float32)
            th = np.float32(np.deg2rad(37.0)); cs, sn = np.float32(np.cos(th)), np.float32(np.sin(th))
            rt = np.array([cs, -sn, 0, 3.5, sn, cs, 0, -1.25, 0, 0, 1, 0.2], np.float32)
            t0 = time.perf_counter()
            tx, ty, tz = O.transform(rt, cx, cy, cz); O.save_as_mat(cx, cy, cz); O.save_as_mat(tx, ty, tz)
            cpu["cloud_manip_2M_points_port_ms_1_core"] = (time.perf_counter() - t0) * 1e3
        except Exception as e:
            cpu["cloud_manip_2M_points_port_ms_1_core"] = repr(e)[:200]
        if not args.no_cli and os.path.exists(pkg.CLI_PATH):
            try:
                cli = cli_numbers(pkg, synth, O, args)
            except Exception as e:      # informational leg: never lose the bench line over it
                cli = {"error": repr(e)[:300]}

    if rank == 0:
        lim = max(h2d / 1e9 / (p_wall / args.steps * 1e-3), 1e-9)
        line = {"metric": METRIC, "value": value, "unit": "frames/s", "n_gpus": world, "steps": args.steps, "warmup": max(args.warmup, 3),
                "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                "dtype": "f32", "data": "synthetic (%d distinct seeded %s frames per rank, tiled to %d)" % (args.distinct, args.sensor, F),
                "config": workload_config(args, world, n_total, in_bytes),
                "e2e": {"value": e2e_v, "unit": "frames/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                        "frames_per_step_per_gpu": Fe, "ms_per_step": e_wall / args.steps, "api": "bevgen_process_host_compact",
                        "input_staging": "pinned, write-combined" if args.wc else "pinned",
                        "full_layout": {"value": e2e_full_v, "api": "bevgen_process_host", "h2d_bytes_per_step": h2d_full,
                                        "d2h_bytes_per_step": d2h_full, "ms_per_step": f_wall / args.steps},
                        "pcie_alone": {"ms_per_step": p_wall / args.steps, "h2d_GBps": lim, "frames_per_s_if_copy_bound": world * Fe / (p_wall / args.steps * 1e-3)},
                        "limiter": "PCIe host-to-device copy: the e2e step takes %.2f ms, the copy engines alone need %.2f ms for the same %d MB in / %d MB out"
                                   % (e_wall / args.steps, p_wall / args.steps, h2d >> 20, d2h >> 20)},
                "gpu_launches": int(launches), "clocks": clocks, "roofline": roofline, "wall_ms_per_step": wall_ms / args.steps,
                "per_rank": per_rank,
                "rank_spread": {"slowest_rank": slow["rank"], "slowest_ms_per_step": slow["ms_per_step"], "fastest_ms_per_step": fast["ms_per_step"],
                                "sum_of_ranks_frames_per_s": sum(F / (r["ms_per_step"] * 1e-3) for r in per_rank),
                                "note": "value is the max-over-ranks number the contract asks for; identical work per rank, so a spread is the box (see per_rank clocks / stages)"}}
        if parity is not None:
            ok = all(v["ok"] for v in parity.values())
            line["parity_checked"] = {"frames": sum(v["frames"] for v in parity.values()), "ok": ok, **parity}
        if cpu:
            line["cpu_baseline"] = cpu
        if cli:
            line["cli"] = cli
        print(json.dumps(line), flush=True)
    g.close()
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
