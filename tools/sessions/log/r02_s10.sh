#!/bin/bash
# round 2, session 10: the persistent tail path tracer - GPU suite (bit-exact images), then C1 / C2 frame times against the
# hand-over threshold (CRT_TAIL) and the shading batch.
mkdir -p gpurun_out
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 )
for tail in 131072 262144 524288 1048576; do
  echo "== CRT_TAIL=$tail"
  CRT_TAIL=$tail python tools/c1_timeline.py cornell-box 2>&1 | tail -2
  CRT_TAIL=$tail python tools/c1_timeline.py veach-mis 2>&1 | tail -1
done
for lib in variants/libcrt_sb1.so variants/libcrt_sb4.so variants/libcrt_sb16.so; do
  echo "== $lib CRT_TAIL=524288"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_TAIL=524288 python tools/c1_timeline.py cornell-box 2>&1 | tail -1
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_TAIL=524288 python tools/c1_timeline.py veach-mis 2>&1 | tail -1
done
echo "== timeline CRT_TAIL=524288"
CRT_TAIL=524288 CRT_TIMELINE=1 python tools/c1_timeline.py cornell-box 2>&1 | tail -14 | grep -v "+    [0-9]\." | cut -c1-60
