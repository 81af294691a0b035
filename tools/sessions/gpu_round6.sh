#!/bin/bash
mkdir -p gpurun_out
V=cudaraytracing_b200/variants
for v in "$@"; do
  echo "== variant $v"
  CRT_LIB=$V/libcrt_$v.so timeout 600 python tools/queue_check.py 2>&1 | grep -E "random|render"
  CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/pool_sweep.py 1048576
  CRT_LIB=$V/libcrt_$v.so timeout 600 ncu --metrics smsp__inst_executed.sum,smsp__thread_inst_executed.sum,gpu__time_duration.sum,smsp__issue_active.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"k_extend|k_shadow" -s 20 -c 2 --csv python tools/pool_sweep.py 1048576 2>/dev/null | grep -E "k_extend|k_shadow" | awk -F'","' '{print $5, $(NF-2), $NF}' | sed 's/(crt::SceneView.*)//'
done 2>&1 | tee gpurun_out/variants6.log
