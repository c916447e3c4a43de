#!/bin/bash
# tools/kbench.py against several built libraries:  kbench_libs.sh "blur" lib1.so lib2.so ...
what=$1; shift
for lib in "$@"; do
  echo "== $lib"
  python - "$lib" $what <<'PY'
import os, sys, runpy
root = os.getcwd()
sys.path.insert(0, os.path.join(root, "video-watermarking-forgery-detection_b200"))
import wmattack._lib as L
L.LIB_PATH = os.path.abspath(sys.argv[1])
sys.argv = ["kbench.py"] + sys.argv[2:]
runpy.run_path(os.path.join(root, "tools", "kbench.py"), run_name="__main__")
PY
done
