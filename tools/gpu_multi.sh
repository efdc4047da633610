#!/bin/bash
# multi-GPU check: N = $1 ranks under torchrun (our arm + the reference arm), as the driver launches them
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt
timeout 600 python -m pytest tests -m gpu -x -q -k "projection or packed or bird" > gpurun_out/pytest_gpu_multi.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2500 gpurun_out/bench_n$N.json; tail -5 gpurun_out/bench_n$N.err
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/bench_ref_n$N.json 2>> gpurun_out/bench_n$N.err
tail -c 600 gpurun_out/bench_ref_n$N.json
# the C++ CLI sharding frames over the GPUs of the box
python - <<'PY'
import os, sys, subprocess, time, importlib
sys.path.insert(0, os.getcwd())
from _load_pkg import load_pkg, load_synth
pkg, synth = load_pkg(), load_synth()
pcd = importlib.import_module("pcpt_b200.pcd")
root = "/tmp/kf_multi"; os.makedirs(root + "/keyframe_point_cloud", exist_ok=True)
n = 96
for i in range(n):
    pcd.write("%s/keyframe_point_cloud/%06d.pcd" % (root, i), synth.make_frame("HDL_64E", 100 + (i % 8)))
open(root + "/keyframe_pose.csv", "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n, seed=5, step=9.0))) + "\n")
for g in (1, int(os.environ.get("NGPU", "2"))):
    t = time.time()
    r = subprocess.run([pkg.CLI_PATH, root, "HDL_64E", "--gpus", str(g), "--batch", "16", "--json-metrics", "gpurun_out/cli_g%d.json" % g], capture_output=True, text=True)
    print("CLI gpus", g, "rc", r.returncode, "%.2f s" % (time.time() - t), open("gpurun_out/cli_g%d.json" % g).read().strip() if r.returncode == 0 else r.stderr[-500:])
PY
