#!/bin/bash
# round 2, session 9: small queues - blocks that stay per launch (rays per lane target 1 = all / 4 / 8 / 16 / 32) on the
# shipped C1 / C2 frames and at steady state.
mkdir -p gpurun_out
for lib in variants/libcrt_rpl0.so variants/libcrt_rpl4.so libcrt.so variants/libcrt_rpl16.so variants/libcrt_rpl32.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib python tools/c1_timeline.py cornell-box 2>&1 | tail -2
  CRT_LIB=$PWD/cudaraytracing_b200/$lib python tools/c1_timeline.py veach-mis 2>&1 | tail -1
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_TIMELINE=1 python tools/c1_timeline.py cornell-box 2>&1 | tail -27 | grep -v "prepare\|generate\|+    [0-9]\." | cut -c1-60 | tr '\n' ';'; echo
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_SPP=128 QB_NO_BATCH=1 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1 | cut -c1-170
done
