#!/bin/bash
# round-2 session 6: parity suite on the library with the Johannsen-Psaltis Euler-Lagrange RHS and the Morris-Thorne metric;
# kernel timings of the error-scale / K-form variants (Kerr C2) and of the two JP right-hand sides (C5)
mkdir -p gpurun_out
L=gpurun_out/r02_run6.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 -x ) > gpurun_out/r02_pytest_gpu_6.log 2>&1; echo "pytest rc=$?" >> $L
for v in base nr1 nr2 nr3 kf nr2kf nr2w14; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 kerr >> $L 2>&1; done
for v in jpgen base nr2; do GB200_LIB=$PWD/variants/libgradus_b200_$v.so python tools/time_variants.py 2048 jp >> $L 2>&1; done
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_6.log | tail -3
cat $L
