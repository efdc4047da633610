#!/bin/bash
# Round-2 gpurun recipe.  STAGES selects what runs (default: all):  test sweep bench san ncu extra
#   /usr/local/graft/bin/gpurun --timeout 1800 -- 'STAGES="test bench" bash tools/gpu_r2.sh'
STAGES=${STAGES:-"test sweep bench san ncu extra"}
mkdir -p gpurun_out
has() { [[ " $STAGES " == *" $1 "* ]]; }
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
nproc > gpurun_out/host.txt; free -g >> gpurun_out/host.txt; (numactl -H || lscpu | grep -i numa) >> gpurun_out/host.txt 2>&1
if has test; then
  timeout 1500 python -m pytest tests -m gpu -q --durations=12 ${PYTEST_ARGS:-} > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
  tail -25 gpurun_out/pytest_gpu.log
fi
if has sweep; then
  : > gpurun_out/sweep.txt
  for wpb in ${WPBS:-1 2 4}; do
    BEVGEN_FOLD_WPB=$wpb timeout 300 python bench.py --steps 4 --warmup 3 --e2e-frames 48 --no-cpu-baseline --no-cli --no-parity > gpurun_out/sw.json 2>> gpurun_out/sweep.err
    python - "fold_wpb=$wpb" >> gpurun_out/sweep.txt <<'PY'
import json,sys
try:
    d=json.load(open('gpurun_out/sw.json'))
    print(sys.argv[1], "value %.0f frames/s  %.3f us/frame" % (d["value"], 1e6/d["value"]), {k: round(v,3) for k,v in d["roofline"]["stage_us_per_frame"].items()}, "e2e %.0f" % d["e2e"]["value"])
except Exception as e:
    print(sys.argv[1], "FAILED", e)
PY
  done
  cat gpurun_out/sweep.txt
fi
if has bench; then
  timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 6000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
  timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
fi
if has san; then
  # compute-sanitizer over the small-sensor parity tests (memcheck: out-of-bounds / misaligned / leaks; racecheck: shared-memory hazards)
  SEL='(synthetic_frames_bit_exact and HDL_32E) or (random_unstructured and HDL_32E) or (compact_host_path and HDL_32E) or (packed_records and pcd26) or top_flatten or labels_and_major or sweep_fallback or boundaries or above_the_first_segment or cloud_manip_contention'
  timeout 1500 compute-sanitizer --tool memcheck --leak-check full --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "$SEL" > gpurun_out/san_memcheck.log 2>&1; echo "memcheck exit $?" >> gpurun_out/san_memcheck.log
  grep -E "ERROR SUMMARY|LEAK SUMMARY|passed|failed|exit" gpurun_out/san_memcheck.log | tail -6
  timeout 1500 compute-sanitizer --tool racecheck --racecheck-report all --error-exitcode 9 python -m pytest tests/test_gpu_parity.py -m gpu -q -x -k "(synthetic_frames_bit_exact and HDL_32E) or (compact_host_path and HDL_32E) or top_flatten or above_the_first_segment or cloud_manip_contention" > gpurun_out/san_racecheck.log 2>&1; echo "racecheck exit $?" >> gpurun_out/san_racecheck.log
  grep -E "RACECHECK SUMMARY|hazard|passed|failed|exit" gpurun_out/san_racecheck.log | tail -8
fi
if has ncu; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --frames 740 --wave 740 --e2e-frames 8 --no-cpu-baseline --no-cli --no-parity > gpurun_out/ncu_launch.log 2>&1
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(order|ground|sector|seg|finalize)" -s 21 -c 7 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --frames 740 --wave 740 --e2e-frames 8 --no-cpu-baseline --no-cli --no-parity > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log
fi
if has extra; then
  timeout 900 python tools/bench_extra.py > gpurun_out/bench_extra.jsonl 2> gpurun_out/bench_extra.err; cat gpurun_out/bench_extra.jsonl; tail -3 gpurun_out/bench_extra.err
fi
ls -la gpurun_out | head -40
