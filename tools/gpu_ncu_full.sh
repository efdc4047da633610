#!/bin/bash
# usage: KREGEX=k_seg_fold SKIP=0 COUNT=1 bash tools/gpu_ncu_full.sh
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:"${KREGEX}" -s ${SKIP:-3} -c ${COUNT:-1} -f -o gpurun_out/prof_${TAG:-k} python bench.py --steps 1 --warmup 3 --frames ${NCU_FRAMES:-740} --wave ${NCU_FRAMES:-740} --e2e-frames 8 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -3 gpurun_out/ncu_full.log; ls -la gpurun_out/*.ncu-rep
