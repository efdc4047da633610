#!/bin/bash
# one gpurun call: parity tests, bench (+ reference arm), small (streams, wave) sweep, launch list, ncu full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -14 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -c 3000 gpurun_out/bench.json
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/bench_ref.json 2>> gpurun_out/bench.err; cat gpurun_out/bench_ref.json
: > gpurun_out/sweep.txt
for cfg in ${SWEEP:-"3 1480" "4 1110"}; do
  set -- $cfg
  BEVGEN_STREAMS=$1 timeout 300 python bench.py --frames 4440 --wave $2 --steps 4 --warmup 3 --e2e-frames 8 --no-cpu-baseline > gpurun_out/sw.json 2>> gpurun_out/sweep.err
  python - "$1" "$2" >> gpurun_out/sweep.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/sw.json'))
print("streams", sys.argv[1], "wave", sys.argv[2], "value %.0f frames/s  %.3f us/frame" % (d["value"], 1e6/d["value"]), {k: round(v/4.44,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
cat gpurun_out/sweep.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --frames 740 --wave 740 --e2e-frames 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_(order|ground|sector|seg|finalize)" -s 21 -c 7 -f -o gpurun_out/prof python bench.py --steps 1 --warmup 3 --frames 740 --wave 740 --e2e-frames 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
ls -la gpurun_out
