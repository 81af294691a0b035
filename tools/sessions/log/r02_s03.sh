#!/bin/bash
# round 2, session 3: how a warp takes ray indices from the queue counter (CRT_FETCH 0 / 2 / 3, chunk 64 / 128 / 256)
# after folding the two leaf-queue traversals into one template; GPU suite on the default build.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4
for lib in variants/libcrt_r1base.so variants/libcrt_f0.so variants/libcrt_f2.so libcrt.so variants/libcrt_f3c64.so variants/libcrt_f3c256.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_SPP=128 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -3
done
echo "== veach, pair nodes"
for lib in variants/libcrt_r1base.so libcrt.so; do
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=veach-mis QB_SPP=64 QB_NO_BATCH=1 timeout 600 python tools/quick_bench.py ploc8 ploc 2>&1 | tail -4
done
