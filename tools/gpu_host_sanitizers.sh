#!/bin/bash
# Host sanitizers over the CLI's pipeline (thread pool, pinned staging, encoders) on a small keyframe folder.
#   make -C point-cloud-preprocessing-tools_b200 bin/batch_multi_bev_gen_asan bin/batch_multi_bev_gen_tsan   (distro g++)
mkdir -p gpurun_out
python - <<'PY'
import importlib, os, sys
sys.path.insert(0, os.getcwd())
from _load_pkg import load_pkg, load_synth
pkg, synth = load_pkg(), load_synth()
pcd = importlib.import_module("pcpt_b200.pcd")
root = "/dev/shm/kf_san"; os.makedirs(root + "/keyframe_point_cloud", exist_ok=True)
n = 40
for i in range(n):
    pcd.write("%s/keyframe_point_cloud/%06d.pcd" % (root, i), synth.make_frame("HDL_32E", 300 + i))
open(root + "/keyframe_pose.csv", "w").write("\n".join(synth.pose_csv_lines(synth.make_poses(n, seed=5, step=9.0))) + "\n")
PY
B=point-cloud-preprocessing-tools_b200/bin
ASAN_OPTIONS=protect_shadow_gap=0:detect_leaks=0:abort_on_error=0 UBSAN_OPTIONS=print_stacktrace=1 timeout 600 $B/batch_multi_bev_gen_asan /dev/shm/kf_san HDL_32E --batch 8 --threads 8 > gpurun_out/host_asan.out 2> gpurun_out/host_asan.err; echo "asan+ubsan rc $?" | tee -a gpurun_out/host_asan.err
grep -cE "ERROR: AddressSanitizer|runtime error" gpurun_out/host_asan.err | sed 's/^/asan+ubsan reports: /'
tail -2 gpurun_out/host_asan.out
TSAN_OPTIONS=report_signal_unsafe=0:history_size=4 timeout 600 $B/batch_multi_bev_gen_tsan /dev/shm/kf_san HDL_32E --batch 8 --threads 8 > gpurun_out/host_tsan.out 2> gpurun_out/host_tsan.err; echo "tsan rc $?" | tee -a gpurun_out/host_tsan.err
grep -c "WARNING: ThreadSanitizer" gpurun_out/host_tsan.err | sed 's/^/tsan warnings: /'
grep -A12 "WARNING: ThreadSanitizer" gpurun_out/host_tsan.err | grep -E "WARNING|#0|#1|#2" | head -24
tail -2 gpurun_out/host_tsan.out
rm -rf /dev/shm/kf_san
