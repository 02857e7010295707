#!/bin/bash
# round-2 session 10: the forward-mode kernel (K4) -- transfer-function timings with the Kerr-only instantiation, one full ncu
# capture of gb200_dual_kernel<1, Kerr>; target-solver tests
mkdir -p gpurun_out
L=gpurun_out/r02_run10.log
nvidia-smi -L > $L 2>&1
python -m pytest tests/test_target_solver.py tests/test_gpu_dual.py tests/test_transfer_functions.py -m gpu -q >> $L 2>&1
python tools/time_transfer.py >> $L 2>&1
ncu --set full --clock-control none --import-source on -k regex:gb200_dual_kernel -s 20 -c 1 -o gpurun_out/prof_dual_v24 -f python tools/time_transfer.py > gpurun_out/r02_ncu_dual.log 2>&1
cat $L
