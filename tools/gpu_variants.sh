#!/bin/bash
# A/B of compile-time variants on one box:  tools/gpu_variants.sh base <name> ...  where lib/libbevgen_cuda_<name>.so was
# built beforehand with e.g. `nvcc ... -DSCAT_T=64 -shared -o lib/libbevgen_cuda_scat64.so csrc/bevgen_capi.cu`
# (tunables: SCAT_T, SCAT_PPT, SCAT_PRED_LOADS, ORD_THREADS, ORD_PREFETCH, ORD_MIN_CTAS, FOLD_DIST_N, FOLD_SINGLE_BUFFER, FOLD_RAW_DESC,
# SEG_MIN_CTAS).  "base" = the library `make` built.
mkdir -p gpurun_out; L=point-cloud-preprocessing-tools_b200/lib
cp $L/libbevgen_cuda.so /tmp/base.so
for v in "$@"; do
  if [ "$v" = base ]; then cp /tmp/base.so $L/libbevgen_cuda.so; else cp $L/libbevgen_cuda_$v.so $L/libbevgen_cuda.so; fi
  timeout 300 python bench.py --steps 5 --warmup 3 --e2e-frames 8 --no-cpu-baseline --no-cli --no-parity > gpurun_out/b.json 2>gpurun_out/b.err
  python -c "
import json,sys; d=json.load(open('gpurun_out/b.json')); print(sys.argv[1], round(d['value']), {k: round(v/4.44,3) for k,v in d['roofline']['stage_ms_per_step'].items()})" $v
done
cp /tmp/base.so $L/libbevgen_cuda.so
