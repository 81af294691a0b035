#!/bin/bash
# round 2, session 1: wide-node step with the folded bias and the parallel-axis test out of line (libcrt.so) against
# the round-1 binary (variants/libcrt_r1base.so): GPU suite, stage times, instruction counts of k_shadow / k_extend.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1500 python -m pytest tests -m gpu -x -q ) 2>&1 | tail -5
for lib in variants/libcrt_r1base.so libcrt.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -4
done
M=smsp__inst_executed.sum,smsp__thread_inst_executed.sum,sm__inst_executed_pipe_alu.sum,sm__inst_executed_pipe_fma.sum,sm__inst_executed_pipe_fmaheavy.sum,sm__inst_executed_pipe_fmalite.sum,sm__inst_executed_pipe_lsu.sum,gpu__time_duration.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__inst_issued.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active
for lib in variants/libcrt_r1base.so libcrt.so; do
  tag=$(basename $lib .so)
  CRT_LIB=$PWD/cudaraytracing_b200/$lib QB_SCENES=cornell-box QB_NO_BATCH=1 timeout 900 ncu --metrics $M --clock-control none -k regex:'k_shadow|k_extend' -s 4 -c 4 --csv --log-file gpurun_out/r02_s01_inst_$tag.csv python tools/quick_bench.py ploc8 > gpurun_out/ncu_inst_$tag.log 2>&1
  tail -2 gpurun_out/ncu_inst_$tag.log | cut -c1-150
done
