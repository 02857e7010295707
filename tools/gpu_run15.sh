#!/bin/bash
# round-2 session 15 (eight GPUs): torchrun bench of kernel v26 at N = 8, 4, 2 (weak + strong objects, NCCL-reduced line profile)
mkdir -p gpurun_out
L=gpurun_out/r02_run15.log
nvidia-smi -L > $L 2>&1
for N in 8 4 2; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2954$N bench.py --gpus $N --steps 5 --warmup 3 --no-callers > gpurun_out/r02_bench_v26_n$N.json 2> gpurun_out/r02_bench_v26_n$N.err; echo "bench N=$N rc=$?" >> $L
done
cat $L
