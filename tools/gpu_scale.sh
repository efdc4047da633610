#!/bin/bash
# scaling check: bench.py at N ranks under torchrun, as the driver launches it
N=${1:-8}
mkdir -p gpurun_out
free -g | head -2 > gpurun_out/host_mem.txt; nproc >> gpurun_out/host_mem.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 2200 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err; cat gpurun_out/host_mem.txt
