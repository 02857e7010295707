#!/bin/bash
# round-2 session 9 (two GPUs): full parity suite (target solver, staged multi-device render), in-process communicator timings
mkdir -p gpurun_out
L=gpurun_out/r02_run9.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_9.log 2>&1; echo "pytest rc=$?" >> $L
python tools/time_comm.py >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_9.log | tail -3
cat $L
