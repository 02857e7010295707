#!/bin/bash
# round-2 session 5 (two GPUs): full parity suite incl. the two-device tests, in-process multi-GPU timings, caller timings
mkdir -p gpurun_out
L=gpurun_out/r02_run5.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 -s ) > gpurun_out/r02_pytest_gpu_5.log 2>&1; echo "pytest rc=$?" >> $L
python tools/time_comm.py >> $L 2>&1
python tools/time_transfer.py >> $L 2>&1
python tools/time_tf_table.py >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_5.log | tail -3
cat $L
