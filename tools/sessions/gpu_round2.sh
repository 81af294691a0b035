#!/bin/bash
mkdir -p gpurun_out
for w in c1 c2 c3; do
  ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_$w.log 2>&1
done
for w in c1 c2 c3; do
  ( time timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_ref_$w.log 2>&1
done
# launch list of the default bench command (first 400 launches after the BVH build)
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_c3_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_c3_list.log 2>&1
# full capture: k_extend / k_shadow / k_shade in steady state of C3
CRT_CPU_BUDGET=1e5 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shadow|k_shade_compat" -s 30 -c 3 -o gpurun_out/r01_c3_full python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_c3_full.log 2>&1
# full capture: C5 closest / any
CRT_C5_RAYS=20000000 timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_trace_batch" -c 2 -o gpurun_out/r01_c5_full python bench.py --workload c5 --steps 1 --warmup 0 > gpurun_out/ncu_c5_full.log 2>&1
grep -h '^{' gpurun_out/bench_*.log | cut -c1-300
ls -la gpurun_out
