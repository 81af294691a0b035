#!/bin/bash
mkdir -p gpurun_out
V=cudaraytracing_b200/variants
for v in "$@"; do
  echo "== variant $v"
  CRT_LIB=$V/libcrt_$v.so timeout 600 python tools/queue_check.py
  CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/pool_sweep.py 1048576 8388608
  CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/tail_sweep.py 131072
  CRT_LIB=$V/libcrt_$v.so CRT_CPU_BUDGET=1e5 CRT_C5_RAYS=50000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('c5 closest %.1f Mrays/s any %.1f Mrays/s' % (d['value'], d['any_hit']['mrays_s']))
"
done 2>&1 | tee gpurun_out/variants.log
