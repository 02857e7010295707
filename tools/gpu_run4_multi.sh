#!/bin/bash
# round-2 multi-GPU session (gpurun --gpus N): the library's in-process communicator on real distinct GPUs, and the torchrun bench
N=${1:-2}
mkdir -p gpurun_out
L=gpurun_out/r02_run4_n$N.log
nvidia-smi -L > $L 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "comm_entry_points" -s >> $L 2>&1
python tools/time_comm.py >> $L 2>&1
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus $N --steps 5 --warmup 3 --no-callers > gpurun_out/r02_bench_n$N.json 2> gpurun_out/r02_bench_n$N.err; echo "bench rc=$?" >> $L
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29534 bench.py --impl reference --gpus $N --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_n$N.json 2>> $L; echo "ref rc=$?" >> $L
cat $L
