#!/bin/bash
# 4-rank bench:  gpurun --gpus 4 --timeout 900 -- 'bash tools/gpu_check4.sh'
mkdir -p gpurun_out
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29541 \
  bench.py --gpus 4 --steps 10 --warmup 4 --no-cpu-baseline > gpurun_out/bench4.json 2> gpurun_out/bench4.err; echo "bench4 rc=$?"
tail -c 200 gpurun_out/bench4.err; wc -l gpurun_out/bench4.json
