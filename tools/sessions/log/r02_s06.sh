#!/bin/bash
# round 2, session 6: tile-ordered path generation (accumulation buffer of the paths in flight L2-resident at 4K) with
# the per-refill fetch (f0) and the chunked one (default build); shared-memory diet of the wide kernels (steps / stack).
mkdir -p gpurun_out
q() { env CRT_LIB=$PWD/cudaraytracing_b200/$1 QB_SCENES=cornell-box QB_NO_BATCH=1 "${@:2}" timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1; }
for lib in variants/libcrt_f0.so libcrt.so; do
  echo "== $lib 4K tile off / 2^21 / 2^20 / 2^22"
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=0
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=1048576
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=4194304
done
for lib in variants/libcrt_f0s1.so variants/libcrt_f0ws2.so variants/libcrt_f0lean.so variants/libcrt_f0ws0.so; do
  echo "== $lib 1080p spp128, 4K spp48"
  q $lib QB_SPP=128
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48
done
( time timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
