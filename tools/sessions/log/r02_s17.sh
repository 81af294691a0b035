#!/bin/bash
# round 2, session 17: the vertex arithmetic with fewer divisions (shadow-test values keep the reference's sequence) against the
# previous statement in one session; GPU suite; converged-image study.
mkdir -p gpurun_out
for lib in variants/libcrt_prediet.so libcrt.so variants/libcrt_prediet.so libcrt.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_NO_BATCH=1 QB_SPP=128 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -4 | cut -c1-170
done
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
python tools/c1_timeline.py cornell-box 2>&1 | tail -1
python tools/c1_timeline.py veach-mis 2>&1 | tail -1
timeout 1500 python tools/converged_study.py 256 1024 4096 16384 2>&1 | tail -16
