#!/bin/bash
# round 2, session 4: where do the warps of k_shadow wait when the reservation atomic is 9x rarer (CRT_FETCH 2)?
mkdir -p gpurun_out
CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_f2.so QB_SCENES=cornell-box QB_SPP=64 QB_NO_BATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shadow' -s 3 -c 1 -o gpurun_out/r02_s04_f2 -f python tools/quick_bench.py ploc8 > gpurun_out/ncu_f2.log 2>&1
tail -2 gpurun_out/ncu_f2.log | cut -c1-150
CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_f0.so QB_SCENES=cornell-box QB_SPP=64 QB_NO_BATCH=1 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_shadow' -s 3 -c 1 -o gpurun_out/r02_s04_f0 -f python tools/quick_bench.py ploc8 > gpurun_out/ncu_f0.log 2>&1
tail -2 gpurun_out/ncu_f0.log | cut -c1-150
