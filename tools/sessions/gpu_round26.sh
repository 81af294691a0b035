#!/bin/bash
# session 26 (N GPUs, $1): NCCL path on real GPUs: sharded render == single-GPU render bit for bit, then the scaling lines
N=${1:-2}
mkdir -p gpurun_out
L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517"
( time timeout 600 $L tools/multi_gpu_check.py ) > gpurun_out/multi_gpu_check_n$N.log 2>&1
grep -E "GPUs|Error|error" gpurun_out/multi_gpu_check_n$N.log | head -20
for w in c3 c1 c2 c5; do
  ( time timeout 900 $L bench.py --gpus $N --workload $w --steps 3 --warmup 3 ) > gpurun_out/scale_${w}_n$N.log 2>&1
  grep -h '^{' gpurun_out/scale_${w}_n$N.log | cut -c1-200
  grep real gpurun_out/scale_${w}_n$N.log
done
( time timeout 600 $L bench.py --impl reference --gpus $N --steps 3 --warmup 3 ) > gpurun_out/scale_ref_c3_n$N.log 2>&1
grep -h '^{' gpurun_out/scale_ref_c3_n$N.log | cut -c1-160
