#!/bin/bash
# round 2, session 11: tail path tracer tuning (adaptive shading trigger; resident blocks; stash high-water mark), ncu of k_tail,
# the reference arm on the synthetic scene, the pipelined host-buffer ray batches.
mkdir -p gpurun_out
for lib in libcrt.so variants/libcrt_tb3.so variants/libcrt_tb5.so variants/libcrt_th32.so variants/libcrt_th8.so; do
  echo "== $lib"
  CRT_LIB=$PWD/cudaraytracing_b200/$lib python tools/c1_timeline.py cornell-box 2>&1 | tail -2
  CRT_LIB=$PWD/cudaraytracing_b200/$lib python tools/c1_timeline.py veach-mis 2>&1 | tail -1
done
echo "== timeline"
CRT_TIMELINE=1 python tools/c1_timeline.py cornell-box 2>&1 | tail -14 | grep -v "+    [0-9]\." | cut -c1-60
CRT_TIMELINE=1 python tools/c1_timeline.py veach-mis 2>&1 | tail -20 | grep -v "+    [0-9]\." | cut -c1-60
timeout 600 ncu --set full --clock-control none --import-source on -k regex:'k_tail' -s 3 -c 1 -o gpurun_out/r02_s11_tail -f python tools/c1_timeline.py cornell-box > gpurun_out/ncu_tail.log 2>&1
tail -2 gpurun_out/ncu_tail.log | cut -c1-150
echo "== reference arm c5"
( time timeout 900 python bench.py --impl reference --workload c5 ) 2>&1 | tail -5 | cut -c1-1500
echo "== ours c5"
( time timeout 900 python bench.py --workload c5 --steps 2 --warmup 1 ) 2>&1 | tail -5 | cut -c1-2500
