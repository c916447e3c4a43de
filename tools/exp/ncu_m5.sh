#!/bin/bash
# ncu --set full on chosen variants of the experiment binary (one launch each)
VARS=${1:-"0|23|31"}
M5_WARM=0 M5_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:m5_kernel<\(int\)(${VARS}), \(bool\)1>" -c 4 -f -o gpurun_out/m5exp tools/exp/m5exp 64 512 512 > gpurun_out/m5exp_ncu.log 2>&1
tail -3 gpurun_out/m5exp_ncu.log; ls -la gpurun_out
