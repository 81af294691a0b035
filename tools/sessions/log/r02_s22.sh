#!/bin/bash
# round 2, session 22: 9 / 10 resident blocks of the traversal kernels (56 / 51 registers) against 8 (64); ncu of the real tail launch
mkdir -p gpurun_out
q() { env QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | grep "est 0" | cut -c1-170; }
for lib in libcrt.so variants/libcrt_mb9.so variants/libcrt_mb10.so libcrt.so; do
  echo "== $lib"
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=48 QB_W=3840 QB_H=2160 QB_SCENES=cornell-box
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=128
  CRT_LIB=$PWD/cudaraytracing_b200/$lib CRT_C5_RAYS=40000000 CRT_C5_E2E_RAYS=4000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('C5 closest', d['value'], 'any', d['any_hit']['mrays_s'])"
done
timeout 600 ncu --set full --metrics lts__t_bytes.sum --clock-control none --import-source on -k regex:'k_tail' -s 4 -c 1 -o gpurun_out/r02_final_c1 -f python tools/c1_timeline.py cornell-box > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log | cut -c1-120
