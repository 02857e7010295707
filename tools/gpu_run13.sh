#!/bin/bash
# round-2 session 13: kernel v25 (seventh-stage sine/cosine from the sixth stage's) -- parity suite, timings, bench, sanitizer
mkdir -p gpurun_out
L=gpurun_out/r02_run13.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_13.log 2>&1; echo "pytest rc=$?" >> $L
GB200_LIB=$PWD/gradus.jl_b200/csrc/libgradus_b200.so python tools/time_variants.py 2048 kerr >> $L 2>&1
GB200_LIB=$PWD/gradus.jl_b200/csrc/libgradus_b200.so python tools/time_variants.py 2048 jp >> $L 2>&1
( time python bench.py --steps 5 --warmup 3 ) > gpurun_out/r02_bench_13.json 2> gpurun_out/r02_bench_13.err; echo "bench rc=$?" >> $L
python tools/bench_configs.py >> $L 2>&1
S=gpurun_out/r02_compute_sanitizer.txt
echo 'compute-sanitizer --tool memcheck python -m pytest tests -m gpu -k "c3_line or batched or shakura or table or heights or pipelined or target or dual or morris or bucket2d or fused" (trace kernel v25 with the fused histogram, forward-mode and target kernels, batches, pipelined host output)' > $S
compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "c3_line or batched or shakura or table or heights or pipelined or target or dual or morris or bucket2d or fused" 2>&1 | tail -4 >> $S
echo >> $S
echo 'compute-sanitizer --tool racecheck python -m pytest tests -m gpu -k "c3_line or heights or fused" (per-CTA shared histogram with 128-bit adds, shared-memory stage values)' >> $S
compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k "c3_line or heights or fused" 2>&1 | tail -4 >> $S
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_13.log | tail -3
cat $L; cat $S
