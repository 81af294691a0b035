#!/bin/bash
# session 20: new defaults (ploc8; pair kernels with ballot queue + 256-bit loads + shared stack; compat shade capped at 64
# registers; wide flush 16): full GPU suite, smoke, bench both arms, ncu launch list + full capture with source, variants
mkdir -p gpurun_out
V=$PWD/cudaraytracing_b200/variants
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
echo "== default"; timeout 300 python tools/quick_bench.py ploc ploc8 2>&1 | tee gpurun_out/quick_default.log
for v in shmb10 shmb12 wq8 wq12; do echo "== $v"; CRT_LIB=$V/libcrt_$v.so timeout 300 python tools/quick_bench.py ploc8 2>&1 | grep -v incoh | tee gpurun_out/quick_$v.log; done
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1
grep -h '^{' gpurun_out/bench_ref_c3.log | cut -c1-160
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3_default.log 2>&1
grep -h '^{' gpurun_out/bench_c3_default.log | cut -c1-200
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r01_s20_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-200
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 3 -o gpurun_out/r01_s20_full -f python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200
ls -la gpurun_out
