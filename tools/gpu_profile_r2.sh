#!/bin/bash
# Round-2 evidence pass (one GPU): bench record, ncu launch list of the same command, ncu --set full of the
# kernels the round changed.  Summaries are written as text (the .ncu-rep files are too large to bring back).
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile_r2.sh'
mkdir -p gpurun_out
python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r2.json 2> gpurun_out/bench_r2.err; echo "bench rc=$?"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv \
  --log-file gpurun_out/ncu_r2_launches.csv python bench.py --steps 2 --warmup 4 --no-cpu-baseline > gpurun_out/ncu_launch.log 2>&1
echo "ncu launch list rc=$?"
for spec in median5:median median:median noise:gaussnoise diffjpeg:diffjpeg resize:rb_banded; do
  op=${spec%%:*}; rx=${spec##*:}
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o /tmp/prof_$op \
     python tools/prof_one.py $op > /tmp/prof_$op.log 2>&1
  echo "# ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 python tools/prof_one.py $op   (64x3x512x512 fp32)" > gpurun_out/ncu_r2_$op.txt
  python tools/ncu_blocks.py /tmp/prof_$op.ncu-rep 3 >> gpurun_out/ncu_r2_$op.txt 2>&1
  echo "$op rc=$?"
done
ls -la gpurun_out
