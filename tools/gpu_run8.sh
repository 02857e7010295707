#!/bin/bash
# round-2 session 8 (gpurun --gpus 8): torchrun bench at N = 8 and N = 4 (weak + strong objects, NCCL-reduced line profile),
# the reference arm under torchrun, the library's in-process communicator over all eight devices
mkdir -p gpurun_out
L=gpurun_out/r02_run8.log
nvidia-smi -L > $L 2>&1
python -m pytest tests/test_more_metrics.py -m gpu -q -k "charged_particle_literals or thick_disc_table or morris" >> $L 2>&1
python -m pytest tests/test_gpu_parity.py -m gpu -q -k "comm_entry_points" >> $L 2>&1
python tools/time_comm.py >> $L 2>&1
for N in 8 4; do
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 2953$N bench.py --gpus $N --steps 5 --warmup 3 --no-callers > gpurun_out/r02_bench_v24_n$N.json 2> gpurun_out/r02_bench_v24_n$N.err; echo "bench N=$N rc=$?" >> $L
done
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 bench.py --impl reference --gpus 8 --steps 2 --warmup 1 > gpurun_out/r02_bench_ref_v24_n8.json 2>> $L; echo "ref rc=$?" >> $L
cat $L
