#!/bin/bash
# round-2 session 11 (eight GPUs): the library's in-process communicator with the staged, threaded result scatter
mkdir -p gpurun_out
L=gpurun_out/r02_run11.log
nvidia-smi -L > $L 2>&1
python tools/time_comm.py >> $L 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "comm_entry_points" >> $L 2>&1
cat $L
