#!/bin/bash
# A/B of kernel variants under build/ab/*.so: tools/ab_run.sh TAG config...
TAG=$1; shift
mkdir -p gpurun_out
for lib in build/ab/*.so; do
  RTLSDR_GPU_SCAN_LIB=$PWD/$lib timeout 300 python tools/ab_small.py "$@" 2>&1 | grep -v Warning
done | tee gpurun_out/${TAG}_ab.txt
