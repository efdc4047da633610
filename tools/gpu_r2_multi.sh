#!/bin/bash
# Round-2 multi-GPU recipe:  gpurun --gpus N -- 'bash tools/gpu_r2_multi.sh N'
#   1. the GPU tests that need two devices (CLI --gpus 2 outputs == --gpus 1 outputs, file by file)
#   2. bench.py at N ranks under torchrun, as the driver launches it (+ the same with write-combined input staging)
#   3. BASELINE configs[2], [3]: 10 000-keyframe OS1_64 / HDL_32E batches sharded over the N ranks, label stage at K = 10 000
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/smi_multi.txt; nvidia-smi topo -m >> gpurun_out/smi_multi.txt 2>&1; nproc >> gpurun_out/smi_multi.txt; free -g | head -2 >> gpurun_out/smi_multi.txt
timeout 600 python -m pytest tests -m gpu -q -k "multi_gpu" > gpurun_out/pytest_gpu_multi.log 2>&1; tail -3 gpurun_out/pytest_gpu_multi.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 900 $TR --master-port 29511 bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -c 3000 gpurun_out/bench_n$N.json; tail -3 gpurun_out/bench_n$N.err
timeout 900 $TR --master-port 29513 bench.py --gpus $N --steps 5 --warmup 3 --wc --no-parity > gpurun_out/bench_n${N}_wc.json 2>> gpurun_out/bench_n$N.err
python - $N <<'PY'
import json, sys
n = sys.argv[1]
for tag in ("", "_wc"):
    try:
        d = json.load(open("gpurun_out/bench_n%s%s.json" % (n, tag)))
        print("N=%s%s value %.0f e2e %.0f  per-rank e2e h2d GB/s %s  pcie-alone h2d %s" % (n, tag, d["value"], d["e2e"]["value"],
              [round(r["e2e_h2d_GBps"], 1) for r in d["per_rank"]], [round(r["pcie_alone_h2d_GBps"], 1) for r in d["per_rank"]]))
    except Exception as e:
        print("N=%s%s FAILED %r" % (n, tag, e))
PY
timeout 900 $TR --master-port 29512 tools/bench_extra.py --sharded > gpurun_out/sharded_n$N.jsonl 2> gpurun_out/sharded_n$N.err; cat gpurun_out/sharded_n$N.jsonl; tail -3 gpurun_out/sharded_n$N.err
