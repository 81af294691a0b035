#!/bin/bash
# round 2, session 15: compute-sanitizer (memcheck, initcheck, racecheck); the default bench line of both arms.
bash tools/sessions/r02_sanitizer.sh
echo "== bench reference"
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1; grep -h '^{' gpurun_out/bench_ref_c3.log | cut -c1-600; grep real gpurun_out/bench_ref_c3.log
echo "== bench ours"
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3.log 2>&1; grep -h '^{' gpurun_out/bench_c3.log | cut -c1-600; grep real gpurun_out/bench_c3.log; tail -3 gpurun_out/bench_c3.log | cut -c1-300
