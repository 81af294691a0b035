#!/bin/bash
# session 21: chunked ray-index reservation + lookahead prefetch, shadow contribution in shared memory, plain any-hit store
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
echo "== default (any-hit plain store only)"; timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | tee gpurun_out/quick_default.log
for v in scr ck64 ck64s ck32s ck128s; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | tee gpurun_out/quick_$v.log; done
for v in ck64s; do
  echo "== parity $v"
  ( CRT_LIB=$V/libcrt_$v.so timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_$v.log 2>&1
  tail -4 gpurun_out/pytest_$v.log
done
( timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_wide.py tests/test_ploc.py -m gpu -q ) > gpurun_out/pytest_default.log 2>&1
tail -3 gpurun_out/pytest_default.log
