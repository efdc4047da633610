#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum,smsp__inst_executed.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -s ${NCU_SKIP:-24} -c ${NCU_COUNT:-8} --csv --log-file gpurun_out/launches.csv python bench.py --steps 1 --warmup 3 --frames ${NCU_FRAMES:-740} --wave ${NCU_FRAMES:-740} --e2e-frames 8 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
python - <<'PY'
import csv
rows=[r for r in csv.reader(open('gpurun_out/launches.csv')) if len(r)>14 and r[0].isdigit()]
agg={}
for r in rows:
    agg.setdefault((r[0], r[4].split('(')[0]), {})[r[12]]=r[14]
for (i,k),m in agg.items():
    print(i,k,m)
PY
