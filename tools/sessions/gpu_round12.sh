#!/bin/bash
# session 12: wide-BVH mismatch hunt, variant timings (I2F fix, 256-bit node loads), ncu of the wide kernels
mkdir -p gpurun_out
timeout 900 python tools/debug_wide.py > gpurun_out/debug_wide.log 2>&1; tail -60 gpurun_out/debug_wide.log
echo "== w0 variant"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_w0.so timeout 600 python tools/debug_wide.py 2>&1 | grep -E "differing|mismatches [1-9]" | head -20
echo "== default lib"; timeout 600 python tools/quick_bench.py lbvh lbvh8 2>&1 | tee gpurun_out/quick_default.log
echo "== ld256"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_ld256.so timeout 600 python tools/quick_bench.py lbvh 2>&1 | tee gpurun_out/quick_ld256.log
echo "== w0"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_w0.so timeout 600 python tools/quick_bench.py lbvh8 2>&1 | tee gpurun_out/quick_w0.log
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow' -s 40 -c 4 -o gpurun_out/r01_s12_wide_full -f python bench.py --steps 1 --warmup 0 --builder lbvh8 > gpurun_out/ncu_full_wide.log 2>&1
tail -2 gpurun_out/ncu_full_wide.log | cut -c1-200
