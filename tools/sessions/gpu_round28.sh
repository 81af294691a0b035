#!/bin/bash
# session 28: traversal kernels capped at 64 registers by default: full GPU suite; 51 / 42 registers on cornell-box and C5 / C4
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
export QB_NO_BATCH=1 QB_SPP=128 QB_SCENES=cornell-box
echo "== default"; timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_default_128.log
for v in mb10 mb12; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_${v}_128.log; done
for v in default mb10 mb12; do
  lib=$PWD/cudaraytracing_b200/libcrt.so; [ $v != default ] && lib=$V/libcrt_$v.so
  for w in c5 c4; do for b in ploc8 ploc; do
    echo "== $v $w $b"; CRT_LIB=$lib timeout 600 python bench.py --workload $w --builder $b --steps 3 --warmup 3 2>&1 | grep '^{' | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print(d['value'], d['unit'], d.get('any_hit'), d['roofline']['achieved'], d['roofline'].get('per_ray'))"
  done; done
done
