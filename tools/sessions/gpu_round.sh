#!/bin/bash
# One GPU session: tests, every bench workload, the reference arm, and an ncu launch list. Logs -> gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpu.txt 2>&1
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
for w in c1 c2 c3 c4 c5; do
  ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_$w.log 2>&1
done
for w in c1 c2 c3; do
  ( time timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_ref_$w.log 2>&1
done
tail -3 gpurun_out/pytest_gpu.log
grep -h '^{' gpurun_out/bench_*.log | cut -c1-400
