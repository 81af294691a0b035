#!/bin/bash
mkdir -p gpurun_out
V=cudaraytracing_b200/variants
v=$1
CRT_LIB=$V/libcrt_$v.so timeout 900 ncu --set full --clock-control none --import-source on -k regex:"k_extend|k_shadow" -s 20 -c 2 -o gpurun_out/var_$v python tools/pool_sweep.py 1048576 > gpurun_out/ncu_var_$v.log 2>&1
ls -la gpurun_out/var_$v.ncu-rep
