#!/bin/bash
# session 29: evidence run of the final defaults (traversal kernels at 64 registers): smoke, both arms on C3, launch list,
# full capture of the C3 kernels and of the HBM-bound C5 batch kernel, C1/C2/C4/C5 lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
run() { tag=$1; shift
  ( time timeout 1200 python bench.py "$@" ) > gpurun_out/bench_$tag.log 2>&1
  grep -h '^{' gpurun_out/bench_$tag.log | cut -c1-200; grep real gpurun_out/bench_$tag.log; }
run ref_c3 --impl reference --steps 3 --warmup 3
run c3 --steps 3 --warmup 3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r01_s29_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-120
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 3 -o gpurun_out/r01_s29_full -f python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-120
CRT_C5_RAYS=20000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_trace_batch' -s 1 -c 1 -o gpurun_out/r01_s29_c5 -f python bench.py --workload c5 --steps 1 --warmup 1 > gpurun_out/ncu_c5.log 2>&1
tail -1 gpurun_out/ncu_c5.log | cut -c1-120
run c1 --workload c1 --steps 5 --warmup 3
run c2 --workload c2 --steps 5 --warmup 3
run c5 --workload c5 --steps 3 --warmup 3
run c4 --workload c4 --steps 3 --warmup 3
