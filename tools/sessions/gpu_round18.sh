#!/bin/bash
# session 18: variants of the traversal kernels (ballot-positioned leaf queue, 40-byte triangle reads, 56-byte node reads,
# shared-memory stack, 256-bit node loads) on ploc / ploc8; parity of the combined variants; C3 with ploc8
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
V=$PWD/cudaraytracing_b200/variants
echo "== default"; timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | grep -v incoh | tee gpurun_out/quick_default.log
for v in qb tri40 n364 all5 all6; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | grep -v incoh | tee gpurun_out/quick_$v.log; done
for v in all5 all6; do
  echo "== parity $v"
  ( CRT_LIB=$V/libcrt_$v.so timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_ploc.py tests/test_gpu_wide.py tests/test_gpu_synthetic.py -m gpu -q ) > gpurun_out/pytest_$v.log 2>&1
  tail -4 gpurun_out/pytest_$v.log
done
( time timeout 900 python bench.py --steps 3 --warmup 3 --builder ploc8 ) > gpurun_out/bench_c3_ploc8.log 2>&1
grep -h '^{' gpurun_out/bench_c3_ploc8.log | cut -c1-200
ls -la gpurun_out
