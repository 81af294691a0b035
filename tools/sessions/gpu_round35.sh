#!/bin/bash
# session 35 (short): counters in separate 128-byte lines, refill threshold 8 / 16, pool 2^25
export QB_NO_BATCH=1 QB_SCENES=cornell-box QB_SPP=128
V=$PWD/cudaraytracing_b200/variants
echo "== default"; python tools/quick_bench.py ploc8 2>&1
for v in lines rf8 rf16; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so python tools/quick_bench.py ploc8 2>&1; done
echo "== pool 2^25"; CRT_POOL=33554432 python tools/quick_bench.py ploc8 2>&1
