#!/bin/bash
# round-2 final verification: parity suite, smoke, bench (both arms)
mkdir -p gpurun_out
L=gpurun_out/r02_run22.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_22.log 2>&1; echo "pytest rc=$?" >> $L
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
( time python bench.py --impl reference ) > gpurun_out/r02_bench_ref_22.json 2>> $L
( time python bench.py ) > gpurun_out/r02_bench_22.json 2>> $L; echo "bench rc=$?" >> $L
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_22.log | tail -3
cat $L
