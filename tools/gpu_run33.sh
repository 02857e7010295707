#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run33.log
nvidia-smi -L > $L 2>&1
echo "== default library (12 warps / SM everywhere)" >> $L
python tools/bench_configs.py 2>&1 | grep -E "^C|^ " >> $L
echo "== Kerr instantiations at 16 warps / SM (128 registers)" >> $L
GB200_LIB=$PWD/variants/libgradus_b200_k16.so python tools/bench_configs.py 2>&1 | grep -E "^C|^ " >> $L
cat $L
