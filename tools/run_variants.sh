#!/bin/bash
# time every variants/*.so on the C2 render (one line each) -> gpurun_out/variants.log
mkdir -p gpurun_out
: > gpurun_out/variants.log
for f in ${@:-variants/*.so}; do
  GB200_LIB=$PWD/$f timeout 300 python tools/time_variants.py 2048 >> gpurun_out/variants.log 2>&1
done
cat gpurun_out/variants.log
