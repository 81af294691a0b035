#!/bin/bash
# round 2, session 27: queue-counter atomic issued before the leaf flush and read after it (default) against the atomic per refill (nofa)
mkdir -p gpurun_out
for rep in 1 2; do
for v in variants/libcrt_nofa.so libcrt.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=48 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=4 QB_NO_BATCH=1 QB_SCENES=veach-mis timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v SP_ONLY_UNORDERED=1 timeout 300 python tools/sort_probe.py
done
done 2>&1 | tee gpurun_out/r02_s27.log
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py tests/test_gpu_synthetic.py -m gpu -x -q 2>&1 | tail -3 ) | tee -a gpurun_out/r02_s27.log
