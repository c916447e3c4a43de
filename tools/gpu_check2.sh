#!/bin/bash
# Two-rank pass of both bench arms (what the driver's scaling run does at N=2):
#   gpurun --gpus 2 --timeout 1200 -- 'bash tools/gpu_check2.sh'
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
$TR --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/bench2.json 2> gpurun_out/bench2.err; echo "bench2 rc=$?"
tail -c 400 gpurun_out/bench2.err; wc -l gpurun_out/bench2.json
$TR --master-port 29512 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/ref2.json 2> gpurun_out/ref2.err; echo "ref2 rc=$?"
tail -c 300 gpurun_out/ref2.err; wc -l gpurun_out/ref2.json
