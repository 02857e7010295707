#!/bin/bash
# round-2 GPU session 2: parity suite on the new kernels, bench line, variants, caller timings, launch list
mkdir -p gpurun_out
L=gpurun_out/r02_run2.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 -s ) > gpurun_out/r02_pytest_gpu_2.log 2>&1; echo "pytest rc=$?" >> $L
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_2.json 2> gpurun_out/r02_bench_2.err; echo "bench rc=$?" >> $L
tools/run_variants.sh >> $L 2>&1
GB200_LIB=$PWD/gradus.jl_b200/csrc/libgradus_b200.so python tools/time_variants.py 2048 >> $L 2>&1
python tools/time_transfer.py >> $L 2>&1
python tools/time_tf_table.py >> $L 2>&1
python tools/bench_configs.py >> $L 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_2.log | tail -3
cat $L
