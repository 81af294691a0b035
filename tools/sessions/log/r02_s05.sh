#!/bin/bash
# round 2, session 5: round-robin chunks + dynamic tail + L2 prefetch of the next chunk (CRT_FETCH 4, default build)
# against the per-refill atomic (f0) and the round-1 binary, at 1080p and at 4K; chunk size; leaner shared memory.
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -4
for lib in variants/libcrt_r1base.so variants/libcrt_f0.so libcrt.so variants/libcrt_f4c64.so variants/libcrt_f4c256.so variants/libcrt_lean.so variants/libcrt_steps3.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_SPP=128 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -3
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_W=3840 QB_H=2160 QB_SPP=48 QB_NO_BATCH=1 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1
done
echo "== veach"
for lib in variants/libcrt_r1base.so libcrt.so; do
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=veach-mis QB_SPP=64 QB_NO_BATCH=1 timeout 600 python tools/quick_bench.py ploc8 ploc 2>&1 | tail -4
done
