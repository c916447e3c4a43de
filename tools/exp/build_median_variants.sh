#!/bin/bash
# Builds variant libraries that differ in the median backward's tile configuration (csrc/median.cu, MBCfg):
#   build_median_variants.sh "name:-DWM_MB3_TH=32 -DWM_MB3_STAGES=3 -DWM_MB3_MINB=3" ...
# -> gpurun_variants/libwm_<name>.so (the other objects come from the regular build)
set -e
CS=video-watermarking-forgery-detection_b200/csrc
mkdir -p gpurun_variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $defs -Xptxas -v \
       -c $CS/median.cu -o gpurun_variants/median_$name.o 2> gpurun_variants/median_$name.log
  grep -A1 "median_bwd_tma_kernelILi[35]ELb0" gpurun_variants/median_$name.log | grep -E "registers|spill" | tr '\n' ' '; echo " <- $name"
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/libwm_$name.so gpurun_variants/median_$name.o $(ls $CS/obj/*.o | grep -v /median.o)
done
