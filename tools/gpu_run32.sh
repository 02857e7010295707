#!/bin/bash
mkdir -p gpurun_out
L=gpurun_out/r02_run32.log
nvidia-smi -L > $L 2>&1
for v in base w14 w16 s8 s32 t12 t24 ru10 base; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 kerr >> $L 2>&1; done
for v in base w14 s32 t24; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 jp >> $L 2>&1; done
cat $L
