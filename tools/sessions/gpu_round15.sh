#!/bin/bash
# session 15: full GPU suite, smoke, default bench both arms, ncu launch list, C1/C2 lines, variants, C5/C4
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
summ='
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if "unavailable" in d: print(d); continue
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f launches %s roof %s cpu %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), json.dumps(d.get("roofline"))[:700], d.get("cpu_baseline"), d.get("clocks")))
'
run_w() {  # workload
  w=$1
  ( time timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_ref_$w.log 2>&1
  grep -h '^{' gpurun_out/bench_ref_$w.log | python -c "$summ" "ref $w"
  for b in lbvh lbvh8; do
    ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_${w}_$b.log 2>&1
    grep -h '^{' gpurun_out/bench_${w}_$b.log | python -c "$summ" "$w $b"
    tail -3 gpurun_out/bench_${w}_$b.log | grep real
  done
}
run_w c3
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_s15_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
run_w c1
run_w c2
echo "== default lib"; timeout 600 python tools/quick_bench.py lbvh lbvh8 2>&1 | tee gpurun_out/quick_default.log
for v in ld256 sh6 sh8; do echo "== $v"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_$v.so timeout 600 python tools/quick_bench.py lbvh 2>&1 | tee gpurun_out/quick_$v.log; done
for v in mb6; do echo "== $v"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_$v.so timeout 600 python tools/quick_bench.py lbvh8 2>&1 | tee gpurun_out/quick_$v.log; done
for w in c5 c4; do for b in lbvh lbvh8; do
  ( time timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_${w}_$b.log 2>&1
  grep -h '^{' gpurun_out/bench_${w}_$b.log | python -c "$summ" "$w $b"
  tail -3 gpurun_out/bench_${w}_$b.log | grep real
done; done
ls -la gpurun_out
