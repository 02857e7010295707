#!/bin/bash
# round-2 session 16: kernel v26 -- ncu launch list of the bench command, full captures (Kerr, Johannsen-Psaltis: one launch of the whole image), reference arm
mkdir -p gpurun_out
L=gpurun_out/r02_run16.log
nvidia-smi -L > $L 2>&1
( time python bench.py --impl reference --steps 3 --warmup 1 ) > gpurun_out/r02_bench_ref_v26.json 2>> $L
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02_v26_launches.csv python bench.py --steps 2 --warmup 1 --no-callers --no-strong --no-cpu-baseline > gpurun_out/r02_ncu_bench.log 2>&1
GB200_NO_PIPELINE=1 ncu --set full --clock-control none --import-source on -k regex:gb200_trace_kernel -s 1 -c 1 -o gpurun_out/prof_trace_v26 -f python tools/time_variants.py 2048 kerr > gpurun_out/r02_ncu_full.log 2>&1
GB200_NO_PIPELINE=1 ncu --set full --clock-control none --import-source on -k regex:gb200_trace_kernel -s 1 -c 1 -o gpurun_out/prof_trace_v26_jp -f python tools/time_variants.py 2048 jp >> gpurun_out/r02_ncu_full.log 2>&1
cat $L
