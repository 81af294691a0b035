#!/bin/bash
# round 2, session 31: ray origin + queue index (default, lean1) and also the inverse direction (lean2) kept in the warp's shared struct
# between turns instead of registers, against the registers (lean0)
mkdir -p gpurun_out
for rep in 1 2; do
for v in variants/libcrt_lean0.so libcrt.so variants/libcrt_lean2.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=48 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=4 QB_NO_BATCH=1 QB_SCENES=veach-mis timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v SP_ONLY_UNORDERED=1 timeout 300 python tools/sort_probe.py
done
done 2>&1 | tee gpurun_out/r02_s31.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py -m gpu -x -q 2>&1 | tail -3 ) | tee -a gpurun_out/r02_s31.log
