#!/bin/bash
# round-2 session 17: final state -- parity suite, config table, transfer-function timings
mkdir -p gpurun_out
L=gpurun_out/r02_run17.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_17.log 2>&1; echo "pytest rc=$?" >> $L
python tools/bench_configs.py >> $L 2>&1
python tools/time_transfer.py >> $L 2>&1
python tools/time_tf_table.py >> $L 2>&1
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_17.log | tail -3
cat $L
