#!/bin/bash
# round 2, session 13: C5 regression hunt - round-1 binary, current traversal with the round-1 node step, current build; 40M and 100M rays
mkdir -p gpurun_out
for n in 40000000 100000000; do
for lib in variants/libcrt_r1base.so variants/libcrt_r1node.so libcrt.so; do
  echo "== $lib n=$n"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_C5_RAYS=$n CRT_C5_E2E_RAYS=4000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('closest', d['value'], 'any', d['any_hit']['mrays_s'], 'e2e', d['e2e']['value'])"
done
done
echo "== C1/C2 frame"
python tools/c1_timeline.py cornell-box 2>&1 | tail -1
python tools/c1_timeline.py veach-mis 2>&1 | tail -1
