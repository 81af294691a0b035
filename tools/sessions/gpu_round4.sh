#!/bin/bash
mkdir -p gpurun_out
V=cudaraytracing_b200/variants
for bits in 0 2 3 4 5 6 7; do
  echo "== sort bits $bits"
  CRT_LIB=$V/libcrt_sort.so CRT_EXP_BITS=$bits timeout 300 python tools/pool_sweep.py 1048576 4194304
done 2>&1 | tee gpurun_out/sort_exp.log
