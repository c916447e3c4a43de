#!/bin/bash
# 8-rank bench (what the driver's scaling run does at N=8):  gpurun --gpus 8 --timeout 900 -- 'bash tools/gpu_check8.sh'
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29531 \
  bench.py --gpus 8 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench8.json 2> gpurun_out/bench8.err; echo "bench8 rc=$?"
tail -c 300 gpurun_out/bench8.err; wc -l gpurun_out/bench8.json
