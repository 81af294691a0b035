#!/bin/bash
# round 2, session 18: leaf flush loading two triangles at once (default) against one at a time (nopair); larger leaves (QB_THRESH)
mkdir -p gpurun_out
q() { env QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | grep "est 0" | cut -c1-170; }
for lib in variants/libcrt_nopair.so libcrt.so variants/libcrt_nopair.so libcrt.so; do
  echo "== $lib"
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=128
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=48 QB_W=3840 QB_H=2160 QB_SCENES=cornell-box
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_C5_RAYS=40000000 CRT_C5_E2E_RAYS=4000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('C5 closest', d['value'], 'any', d['any_hit']['mrays_s'])"
done
echo "== leaf size (default build): thresh 2 / 4 / 8"
q QB_SPP=128 QB_SCENES=cornell-box
q QB_SPP=128 QB_SCENES=cornell-box QB_THRESH=4
q QB_SPP=128 QB_SCENES=cornell-box QB_THRESH=8
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
