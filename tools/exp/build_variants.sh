#!/bin/bash
# Builds variant libraries that differ in -D configuration macros of ONE csrc file:
#   build_variants.sh blur "a:-DWM_BT_TH=32 -DWM_BT_STAGES=6 -DWM_BT_MINB=2" "b:..."
# -> gpurun_variants/libwm_<file>_<name>.so (the other objects come from the regular build; run `python -m wmattack.build` first)
set -e
CS=video-watermarking-forgery-detection_b200/csrc
src=$1; shift
mkdir -p gpurun_variants
for spec in "$@"; do
  name=${spec%%:*}; defs=${spec#*:}
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -lineinfo -Xcompiler -fPIC --expt-relaxed-constexpr $defs -Xptxas -v \
       -c $CS/$src.cu -o gpurun_variants/${src}_$name.o 2> gpurun_variants/${src}_$name.log
  echo "$name: $(grep -c 'spill stores, [1-9]' gpurun_variants/${src}_$name.log || true) kernels with spills"
  nvcc -shared -gencode arch=compute_100a,code=sm_100a -o gpurun_variants/libwm_${src}_$name.so gpurun_variants/${src}_$name.o $(ls $CS/obj/*.o | grep -v /$src.o)
done
