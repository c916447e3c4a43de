#!/bin/bash
ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
  -k 'regex:m5b_scatter_kernel<\(int\)35, \(int\)3, \(int\)160>' -c 1 -f -o gpurun_out/m5b tools/exp/m5bexp 64 512 512 > gpurun_out/m5b_ncu.log 2>&1
tail -2 gpurun_out/m5b_ncu.log; ls -la gpurun_out/m5b*
