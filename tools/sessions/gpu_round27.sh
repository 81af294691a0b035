#!/bin/bash
# session 27: occupancy of the wide traversal kernels (launch bounds 7 / 8 blocks per SM), new scheduling-knob test
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
export QB_NO_BATCH=1
for spp in 128 16; do
  export QB_SPP=$spp QB_SCENES=cornell-box
  echo "== default spp $spp"; timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_default_$spp.log
  for v in mb7 mb8; do echo "== $v spp $spp"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_$spp.log; done
done
export QB_SPP=64 QB_SCENES=veach-mis
echo "== default veach"; timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_default_veach.log
for v in mb8; do echo "== $v veach"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_${v}_veach.log; done
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q ) > gpurun_out/pytest_parity.log 2>&1
tail -3 gpurun_out/pytest_parity.log
for b in ploc8; do CRT_LIB=$V/libcrt_mb8.so timeout 600 python bench.py --workload c5 --steps 3 --warmup 3 2>&1 | grep '^{' | cut -c1-150; done
