#!/bin/bash
# session 31 (N GPUs, $1): sharded render == single-GPU render on N GPUs, C3 and C5 scaling lines
N=${1:-4}
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29519"
( time timeout 300 $L tools/multi_gpu_check.py ) > gpurun_out/multi_gpu_check_n$N.log 2>&1
grep -E "GPUs|Error|error" gpurun_out/multi_gpu_check_n$N.log | head -10
for w in c3 c5; do
  ( time timeout 600 $L bench.py --gpus $N --workload $w --steps 3 --warmup 3 ) > gpurun_out/scale_${w}_n$N.log 2>&1
  grep -h '^{' gpurun_out/scale_${w}_n$N.log | cut -c1-200
  grep real gpurun_out/scale_${w}_n$N.log
done
