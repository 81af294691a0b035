#!/bin/bash
# session 19: shade occupancy (launch bounds), pair-node combination (ballot queue + 256-bit loads + shared stack), wide queue knobs
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
echo "== default"; timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | grep -v incoh | tee gpurun_out/quick_default.log
for v in shmb6 shmb8 wmb6 wq16 wq32 ws3 ws1; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | grep -v incoh | tee gpurun_out/quick_$v.log; done
for v in pair3; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc 2>&1 | grep -v incoh | tee gpurun_out/quick_$v.log; done
