#!/bin/bash
# round 2, session 39: with the gate, three node steps per turn (steps3; steps3all: closest-hit batches gated too; gateall: two steps, closest gated)
mkdir -p gpurun_out
for v in libcrt.so variants/libcrt_steps3.so variants/libcrt_steps3all.so variants/libcrt_gateall.so libcrt.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=48 QB_NO_BATCH=1 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=2 QB_NO_BATCH=1 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v SP_ONLY_UNORDERED=1 timeout 300 python tools/sort_probe.py
done 2>&1 | tee gpurun_out/r02_s39.log
