#!/bin/bash
# e2e (host-buffer path) against the staging chunk size:  tools/gpu_chunk.sh 16 24 48 96
mkdir -p gpurun_out
for c in "$@"; do
  BEVGEN_HOST_CHUNK=$c timeout 300 python bench.py --steps 5 --warmup 3 --frames 740 --wave 740 --e2e-frames ${E2E_FRAMES:-256} --no-cpu-baseline --no-cli --no-parity > gpurun_out/ck.json 2>gpurun_out/ck.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/ck.json')); e=d['e2e']; print('chunk', sys.argv[1], 'e2e %.0f frames/s (%.2f ms/step), full layout %.0f, copy engines alone %.0f' % (e['value'], e['ms_per_step'], e['full_layout']['value'], e['pcie_alone']['frames_per_s_if_copy_bound']))" $c
done
