#!/bin/bash
# round 2, session 37: the second node step of a turn only while at least N lanes still have a node in hand (N = 12 / 16 / 20 / 24) against always (default)
mkdir -p gpurun_out
for v in libcrt.so variants/libcrt_ml12.so variants/libcrt_ml16.so variants/libcrt_ml20.so variants/libcrt_ml24.so libcrt.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=48 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=4 QB_NO_BATCH=1 QB_SCENES=veach-mis timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=2 QB_NO_BATCH=1 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v SP_ONLY_UNORDERED=1 timeout 300 python tools/sort_probe.py
done 2>&1 | tee gpurun_out/r02_s37.log
