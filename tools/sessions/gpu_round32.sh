#!/bin/bash
# session 32: wide-node plane bytes decoded through the half unit (CRT_WHALF=1, default) vs PRMT + FADD (variant whalf0)
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
( timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
export QB_NO_BATCH=1 QB_SCENES=cornell-box
for spp in 128 16; do
  export QB_SPP=$spp
  echo "== whalf1 spp $spp"; timeout 200 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_whalf1_$spp.log
  echo "== whalf0 spp $spp"; CRT_LIB=$V/libcrt_whalf0.so timeout 200 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_whalf0_$spp.log
done
for v in 1 0; do
  lib=$PWD/cudaraytracing_b200/libcrt.so; [ $v = 0 ] && lib=$V/libcrt_whalf0.so
  echo "== c5 whalf$v"; CRT_C5_RAYS=40000000 CRT_LIB=$lib timeout 300 python bench.py --workload c5 --steps 3 --warmup 2 2>&1 | grep '^{' | cut -c1-120
done
