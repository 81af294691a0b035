#!/bin/bash
# round 2, session 7: separate the effect of the tile order from the shared-stack depth at 4K (per-refill fetch);
# conditional-graph probe; launch list of the shipped C1 config.
mkdir -p gpurun_out
q() { env CRT_LIB=$PWD/cudaraytracing_b200/$1 QB_SCENES=cornell-box QB_NO_BATCH=1 "${@:2}" timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1 | cut -c1-170; }
for lib in variants/libcrt_f0.so variants/libcrt_f0ws0.so; do
  echo "== $lib 4K tile off / 2^21 ; 1080p"
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=0
  q $lib QB_W=3840 QB_H=2160 QB_SPP=48
  q $lib QB_SPP=128
done
echo "== conditional graph probe"; ./tools/cg_probe_shared
echo "== C1 launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_s07_c1_launches.csv python bench.py --workload c1 --steps 2 --warmup 1 > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log | cut -c1-300
