#!/bin/bash
op=median5; rx=median
timeout 600 ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 -f -o /tmp/prof_$op python tools/prof_one.py $op > /tmp/prof_$op.log 2>&1
echo "# ncu --set full --clock-control none --import-source on -k regex:$rx -s 4 -c 2 python tools/prof_one.py $op   (64x3x512x512 fp32)" > gpurun_out/ncu_r2_$op.txt
python tools/ncu_blocks.py /tmp/prof_$op.ncu-rep 3 >> gpurun_out/ncu_r2_$op.txt 2>&1
echo "$op rc=$?"
