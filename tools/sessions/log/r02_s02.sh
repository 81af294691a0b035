#!/bin/bash
# round 2, session 2: steady-state A/B (spp 128) of the round-1 binary and the new wide-node step, then a full ncu
# capture (with source) of k_extend / k_shade / k_shadow at C3 steady state for the stall breakdown.
mkdir -p gpurun_out
for lib in variants/libcrt_r1base.so libcrt.so variants/libcrt_r1base.so libcrt.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_SPP=128 QB_NO_BATCH=1 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2
done
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 3 -o gpurun_out/r02_s02_full -f python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-120
ls -la gpurun_out/*.ncu-rep
