#!/bin/bash
# round-2 GPU session 3: kernel v23 (Euler-Lagrange Kerr RHS) -- parity suite, bench, occupancy variants, ncu launch list + full capture
mkdir -p gpurun_out
L=gpurun_out/r02_run3.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 -s ) > gpurun_out/r02_pytest_gpu_3.log 2>&1; echo "pytest rc=$?" >> $L
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_3.json 2> gpurun_out/r02_bench_3.err; echo "bench rc=$?" >> $L
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_3.json 2>> $L
tools/run_variants.sh >> $L 2>&1
GB200_LIB=$PWD/gradus.jl_b200/csrc/libgradus_b200.so python tools/time_variants.py 2048 >> $L 2>&1
# launch list of the bench command (cold-cache, serialised: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v23_launches.csv python bench.py --steps 2 --warmup 1 --no-callers --no-strong --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
# one full capture of the trace kernel on the C2 render
ncu --set full --clock-control none --import-source on -k regex:gb200_trace_kernel -s 1 -c 1 -o gpurun_out/prof_trace_v23 -f python tools/time_variants.py 2048 > gpurun_out/r02_ncu_full.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_3.log | tail -3
cat $L
