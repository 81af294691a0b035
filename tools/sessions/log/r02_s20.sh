#!/bin/bash
# round 2, session 20: k_shade at 4K - the vertex arithmetic with fewer divisions (default) against the previous one (prediet), same box
mkdir -p gpurun_out
q() { env QB_NO_BATCH=1 QB_SCENES=cornell-box "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | grep "est 0" | cut -c1-170; }
for lib in variants/libcrt_prediet.so libcrt.so variants/libcrt_prediet.so libcrt.so; do
  echo "== $lib"
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=48 QB_W=3840 QB_H=2160
  q CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SPP=128
done
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 6 -c 1 -o gpurun_out/r02_s20_shade_new -f env QB_NO_BATCH=1 QB_SCENES=cornell-box QB_SPP=48 QB_W=3840 QB_H=2160 python tools/quick_bench.py ploc8 > /dev/null 2>&1
CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_prediet.so timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_shade' -s 6 -c 1 -o gpurun_out/r02_s20_shade_old -f env QB_NO_BATCH=1 QB_SCENES=cornell-box QB_SPP=48 QB_W=3840 QB_H=2160 python tools/quick_bench.py ploc8 > /dev/null 2>&1
ls -la gpurun_out/r02_s20*
