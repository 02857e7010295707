#!/bin/bash
# round-2 GPU session 1: parity suite, bench line, variants, caller timings
mkdir -p gpurun_out
L=gpurun_out/r02_run1.log
nvidia-smi -L > $L 2>&1
nproc >> $L
( time python -m pytest tests -m gpu -q --timeout 900 ) > gpurun_out/r02_pytest_gpu_1.log 2>&1; echo "pytest rc=$?" >> $L
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_1.json 2> gpurun_out/r02_bench_1.err; echo "bench rc=$?" >> $L
tools/run_variants.sh >> $L 2>&1
python tools/time_transfer.py >> $L 2>&1
python tools/time_tf_table.py >> $L 2>&1
( time python bench.py --impl reference --steps 2 --warmup 1 ) > gpurun_out/r02_bench_ref_1.json 2>> $L
tail -5 gpurun_out/r02_pytest_gpu_1.log
cat $L
