#!/bin/bash
# round-2 session 7: kernel v24 (seed-only error-scale reciprocals, Johannsen-Psaltis Euler-Lagrange RHS) -- parity suite,
# bench + reference arm, config table, ncu launch list, full captures of the Kerr and JP kernels (one launch of the whole image)
mkdir -p gpurun_out
L=gpurun_out/r02_run7.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_7.log 2>&1; echo "pytest rc=$?" >> $L
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_7.json 2> gpurun_out/r02_bench_7.err; echo "bench rc=$?" >> $L
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_7.json 2>> $L
python tools/bench_configs.py >> $L 2>&1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v24_launches.csv python bench.py --steps 2 --warmup 1 --no-callers --no-strong --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
GB200_NO_PIPELINE=1 ncu --set full --clock-control none --import-source on -k regex:gb200_trace_kernel -s 1 -c 1 -o gpurun_out/prof_trace_v24 -f python tools/time_variants.py 2048 kerr > gpurun_out/r02_ncu_full.log 2>&1
GB200_NO_PIPELINE=1 ncu --set full --clock-control none --import-source on -k regex:gb200_trace_kernel -s 1 -c 1 -o gpurun_out/prof_trace_v24_jp -f python tools/time_variants.py 2048 jp >> gpurun_out/r02_ncu_full.log 2>&1
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_7.log | tail -3
cat $L
