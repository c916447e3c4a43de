#!/bin/bash
# ncu --set full on chosen template variants of an experiment binary (one launch each):  ncu_exp.sh <binary> <kernel> "<v1|v2>"
BIN=$1; KER=$2; VARS=$3; shift 3
M5_WARM=0 M5_REPS=1 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k "regex:${KER}<\(int\)(${VARS}), \(bool\)1>" -c 4 -f -o gpurun_out/${KER} tools/exp/${BIN} "$@" > gpurun_out/${KER}_ncu.log 2>&1
tail -2 gpurun_out/${KER}_ncu.log; ls -la gpurun_out | head
