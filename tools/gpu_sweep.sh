#!/bin/bash
# sweep of (streams, wave) for the device path
mkdir -p gpurun_out; : > gpurun_out/sweep.txt
for cfg in "1 2220" "2 2220" "2 1110" "3 1480" "4 1110" "4 555" "6 740" "8 555" "8 278"; do
  set -- $cfg
  BEVGEN_STREAMS=$1 timeout 300 python bench.py --frames 4440 --wave $2 --steps 4 --warmup 3 --e2e-frames 8 --no-cpu-baseline > gpurun_out/sw.json 2>> gpurun_out/sweep.err
  python - "$1" "$2" >> gpurun_out/sweep.txt <<'PY'
import json,sys
d=json.load(open('gpurun_out/sw.json'))
print("streams", sys.argv[1], "wave", sys.argv[2], "value %.0f frames/s  %.3f us/frame" % (d["value"], 1e6/d["value"]), {k: round(v/4.44,3) for k,v in d["roofline"]["stage_ms_per_step"].items()})
PY
done
cat gpurun_out/sweep.txt
