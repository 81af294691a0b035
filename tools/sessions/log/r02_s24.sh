#!/bin/bash
# round 2, session 24: k_shade without the slow-path division for zero numerators (default) against the previous arithmetic (nodiv)
# and with the next vertex's hit / triangle record prefetched (pref); parity suite of the default
mkdir -p gpurun_out
for rep in 1 2; do
for v in variants/libcrt_nodiv.so libcrt.so variants/libcrt_pref.so; do
  echo "== $v"
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=3840 QB_H=2160 QB_SPP=48 QB_NO_BATCH=1 QB_SCENES=cornell-box timeout 300 python tools/quick_bench.py ploc8
  env CRT_LIB=$PWD/cudaraytracing_b200/$v QB_W=800 QB_H=600 QB_SPP=4 QB_NO_BATCH=1 QB_SCENES=veach-mis timeout 300 python tools/quick_bench.py ploc8
done
done 2>&1 | tee gpurun_out/r02_s24.log
( timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -3 ) | tee -a gpurun_out/r02_s24.log
