#!/bin/bash
# session 22: variants at steady state (1080p spp 128, cornell-box: pool refilled every iteration like C3) and on short frames
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
export QB_NO_BATCH=1
for spp in 128 16; do
  export QB_SPP=$spp QB_SCENES=cornell-box
  echo "== default spp $spp"; timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_default_$spp.log
  for v in ck64 ck128 tripf wss4 wss8; do echo "== $v spp $spp"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_$spp.log; done
done
export QB_SPP=64 QB_SCENES=veach-mis
echo "== default veach"; timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_default_veach.log
for v in ck64 tripf wss8; do echo "== $v veach"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_veach.log; done
