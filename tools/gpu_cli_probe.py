#!/usr/bin/env python
"""The drop-in CLI on a synthetic HDL_64E keyframe folder: wall numbers and the per-phase CPU time of --json-metrics.
    python tools/gpu_cli_probe.py [n_frames=200] [dir=/dev/shm]"""
import importlib
import json
import os
import shutil
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from _load_pkg import load_pkg, load_synth  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 200
where = sys.argv[2] if len(sys.argv) > 2 else ("/dev/shm" if os.access("/dev/shm", os.W_OK) else None)
pkg, synth = load_pkg(), load_synth()
pcd = importlib.import_module("pcpt_b200.pcd")
base = tempfile.mkdtemp(prefix="bevgen_cli_", dir=where)
try:
    root = os.path.join(base, "kf"); os.makedirs(os.path.join(root, "keyframe_point_cloud"))
    fr = [synth.make_frame("HDL_64E", 5000 + i) for i in range(32)]
    t0 = time.perf_counter()
    for i in range(n):
        pcd.write(os.path.join(root, "keyframe_point_cloud", "%06d.pcd" % i), fr[i % len(fr)])
    open(os.path.join(root, "keyframe_pose.csv"), "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n, seed=5, step=9.0))) + "\n")
    print("folder of %d keyframes written in %.1f s under %s; host threads %d" % (n, time.perf_counter() - t0, base, os.cpu_count()), flush=True)
    cases = (("all files", []), ("all files, again (page cache warm)", []), ("all files, 2 workers per GPU", ["--workers-per-gpu", "2"]),
             ("all files, 3 workers per GPU", ["--workers-per-gpu", "3"]), ("all files, 2 workers, batch 8", ["--workers-per-gpu", "2", "--batch", "8"]),
             ("all files, 4 workers, batch 8", ["--workers-per-gpu", "4", "--batch", "8"]),
             ("no pcd, 2 workers", ["--no-pcd", "--workers-per-gpu", "2"]), ("no encode, no pcd", ["--no-encode", "--no-pcd"]),
             ("no encode, no pcd, 2 workers", ["--no-encode", "--no-pcd", "--workers-per-gpu", "2"]),
             ("all files, batch 32", ["--batch", "32"]), ("all files, batch 32, 2 workers", ["--batch", "32", "--workers-per-gpu", "2"]),
             ("all files, malloc keeps its memory", ["MALLOC_MMAP_THRESHOLD_=268435456", "MALLOC_TRIM_THRESHOLD_=2147483647", "MALLOC_TOP_PAD_=67108864"]),
             ("all files, 24 threads", ["--threads", "24"]), ("all files, 32 threads", ["--threads", "32"]),
             ("all files, 12 threads", ["--threads", "12"]),
             ("all files, malloc keeps its memory, 24 threads", ["MALLOC_MMAP_THRESHOLD_=268435456", "MALLOC_TRIM_THRESHOLD_=2147483647", "MALLOC_TOP_PAD_=67108864", "--threads", "24"]))
    if os.environ.get("CLI_PROBE_CASES"):
        want = set(os.environ["CLI_PROBE_CASES"].split(";"))
        cases = tuple(c for c in cases if c[0] in want)
    for tag, extra in cases:
        mj = os.path.join(base, "m.json")
        t0 = time.perf_counter()
        env = dict(os.environ)
        while extra and "=" in extra[0] and not extra[0].startswith("--"):      # leading NAME=value items: environment of this case
            k, v = extra[0].split("=", 1); env[k] = v; extra = extra[1:]
        r = subprocess.run([pkg.CLI_PATH, root, "HDL_64E", "--json-metrics", mj] + extra, capture_output=True, text=True, timeout=900, env=env)
        wall = time.perf_counter() - t0
        m = json.load(open(mj)) if r.returncode == 0 and os.path.exists(mj) else {"error": r.stderr[-300:]}
        print(json.dumps({"what": tag, "process_wall_s": round(wall, 3), **m}), flush=True)
finally:
    shutil.rmtree(base, ignore_errors=True)
