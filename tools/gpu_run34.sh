#!/bin/bash
# round-2 final: parity suite (incl. README block), smoke, sanitizer on the final kernels
mkdir -p gpurun_out
L=gpurun_out/r02_run34.log
nvidia-smi -L > $L 2>&1
( time python -m pytest tests -m gpu -q --timeout 1200 ) > gpurun_out/r02_pytest_gpu_34.log 2>&1; echo "pytest rc=$?" >> $L
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" >> $L 2>&1
S=gpurun_out/r02_compute_sanitizer.txt
echo 'compute-sanitizer --tool memcheck python -m pytest tests -m gpu -k "c3_line or batched or shakura or table or heights or pipelined or target or dual or morris or bucket2d or fused or polish or dilaton or lag_frequency" (trace kernel v25 with the fused histogram, forward-mode, target and path kernels, batches, pipelined host output)' > $S
compute-sanitizer --tool memcheck python -m pytest tests -m gpu -q -k "c3_line or batched or shakura or table or heights or pipelined or target or dual or morris or bucket2d or fused or polish or dilaton or lag_frequency" 2>&1 | tail -4 >> $S
echo >> $S
echo 'compute-sanitizer --tool racecheck python -m pytest tests -m gpu -k "c3_line or heights or fused" (per-CTA shared histogram with 128-bit adds, shared-memory stage values)' >> $S
compute-sanitizer --tool racecheck python -m pytest tests -m gpu -q -k "c3_line or heights or fused" 2>&1 | tail -4 >> $S
grep -E "passed|failed" gpurun_out/r02_pytest_gpu_34.log | tail -3
cat $L; cat $S
