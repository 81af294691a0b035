#!/bin/bash
# session 13: full GPU suite after the zero-component fix, variant timings, C1/C2/C4/C5 bench lines
mkdir -p gpurun_out
( time timeout 1700 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
echo "== default lib"; timeout 600 python tools/quick_bench.py lbvh lbvh8 2>&1 | tee gpurun_out/quick_default.log
for v in mb6 mb8; do echo "== $v"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_$v.so timeout 600 python tools/quick_bench.py lbvh8 2>&1 | grep cornell | tee gpurun_out/quick_$v.log; done
for v in sh6 sh8 ld256; do echo "== $v"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_$v.so timeout 600 python tools/quick_bench.py lbvh 2>&1 | grep -v incoh | tee gpurun_out/quick_$v.log; done
for w in c1 c2; do
  ( time timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_ref_$w.log 2>&1
  grep -h '^{' gpurun_out/bench_ref_$w.log | cut -c1-120
  for b in lbvh lbvh8; do
    ( time timeout 600 python bench.py --workload $w --steps 5 --warmup 3 --builder $b ) > gpurun_out/bench_${w}_$b.log 2>&1
    grep -h '^{' gpurun_out/bench_${w}_$b.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('$w $b value %.1f ms/step %.3f e2e %.1f launches %d' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
"
  done
done
for b in lbvh lbvh8; do
  ( time timeout 1200 python bench.py --workload c5 --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_c5_$b.log 2>&1
  grep -h '^{' gpurun_out/bench_c5_$b.log | cut -c1-900
  tail -3 gpurun_out/bench_c5_$b.log | grep real
done
for b in lbvh lbvh8; do
  ( time timeout 1200 python bench.py --workload c4 --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_c4_$b.log 2>&1
  grep -h '^{' gpurun_out/bench_c4_$b.log | cut -c1-900
  tail -3 gpurun_out/bench_c4_$b.log | grep real
done
