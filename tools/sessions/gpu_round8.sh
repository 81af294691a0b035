#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log | head -3
V=cudaraytracing_b200/variants
for f in $V/*.so; do
  v=$(basename $f .so)
  echo "== $v"
  CRT_LIB=$f timeout 300 python tools/pool_sweep.py 1048576 8388608
  CRT_LIB=$f CRT_CPU_BUDGET=1e5 CRT_C5_RAYS=50000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 | python -c "
import sys, json
for ln in sys.stdin:
    if ln.startswith('{'):
        d = json.loads(ln); print('c5 closest %.1f Mrays/s any %.1f Mrays/s' % (d['value'], d['any_hit']['mrays_s']))
"
done 2>&1 | tee gpurun_out/variants8.log
