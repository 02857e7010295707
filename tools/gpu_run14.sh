#!/bin/bash
# round-2 session 14: small gb200_trace calls through the staged batch path -- parity suite, config table (C4 lines)
mkdir -p gpurun_out
L=gpurun_out/r02_run14.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_14.log 2>&1; echo "pytest rc=$?" >> $L
python tools/bench_configs.py >> $L 2>&1
GB200_NO_SMALL_BATCH=1 python tools/bench_configs.py 2>&1 | grep C4 | sed 's/^/without the small-call batch path: /' >> $L
python tools/time_corona.py >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_14.log | tail -3
cat $L
