#!/bin/bash
# 2-rank bench of our arm only:  gpurun --gpus 2 --timeout 900 -- 'bash tools/exp/gpu_check2_ours.sh'
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
