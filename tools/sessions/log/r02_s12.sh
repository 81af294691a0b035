#!/bin/bash
# round 2, session 12: C5 regression hunt - round-1 binary against the current one through the same harness
mkdir -p gpurun_out
for lib in variants/libcrt_r1base.so libcrt.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_C5_RAYS=40000000 CRT_C5_E2E_RAYS=4000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('closest', d['value'], 'any', d['any_hit']['mrays_s'], 'e2e', d['e2e']['value'], d['roofline']['per_ray'])"
done
CRT_C5_RAYS=20000000 CRT_C5_E2E_RAYS=1000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_batch' -s 1 -c 1 -o gpurun_out/r02_s12_c5 -f python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-100
