#!/bin/bash
# session 23: tiled pixel order of the camera paths (8x4 tiles, blocks of BxB tiles) vs scanline order; parity of the new default
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
export QB_NO_BATCH=1
for spp in 128 16; do
  export QB_SPP=$spp QB_SCENES=cornell-box
  echo "== default (tile 8x4) spp $spp"; timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_default_$spp.log
  for v in notile tb4 tb8 tb16; do echo "== $v spp $spp"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_$spp.log; done
done
export QB_SPP=64 QB_SCENES=veach-mis
echo "== default veach"; timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_default_veach.log
for v in notile tb8; do echo "== $v veach"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_veach.log; done
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( CRT_LIB=$V/libcrt_tb8.so timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/pytest_tb8.log 2>&1
tail -3 gpurun_out/pytest_tb8.log
