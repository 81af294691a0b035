#!/bin/bash
# round 2, session 21: shadow queue in 8 regions with a counter each (k_shade no longer waits on one address) against the single counter
mkdir -p gpurun_out
q() { env QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | grep "est" | cut -c1-170; }
for lib in variants/libcrt_onecounter.so libcrt.so variants/libcrt_onecounter.so libcrt.so; do
  echo "== $lib"
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=48 QB_W=3840 QB_H=2160 QB_SCENES=cornell-box
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=128
done
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
python tools/c1_timeline.py cornell-box 2>&1 | tail -1
python tools/c1_timeline.py veach-mis 2>&1 | tail -1
