#!/bin/bash
# session 33: final binary (dead experiment code removed): full GPU suite, smoke, one C3 line
mkdir -p gpurun_out
( timeout 600 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
( time timeout 600 python bench.py --steps 2 --warmup 3 ) > gpurun_out/bench_c3_final.log 2>&1
grep -h '^{' gpurun_out/bench_c3_final.log | cut -c1-200
