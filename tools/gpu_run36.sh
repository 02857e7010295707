#!/bin/bash
# round-2: kernel v26 (controller log / exp at the accuracy of their argument) -- parity suite, bench, config table
mkdir -p gpurun_out
L=gpurun_out/r02_run36.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_36.log 2>&1; echo "pytest rc=$?" >> $L
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
( time python bench.py ) > gpurun_out/r02_bench_36.json 2>> $L; echo "bench rc=$?" >> $L
python tools/bench_configs.py >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_36.log | tail -3
cat $L
