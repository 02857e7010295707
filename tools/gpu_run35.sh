#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run35.log
nvidia-smi -L > $L 2>&1
for v in base invq qlo base invq qlo; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 kerr >> $L 2>&1; done
for v in base invq; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 jp >> $L 2>&1; done
cat $L
