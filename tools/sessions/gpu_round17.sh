#!/bin/bash
# session 17: full GPU suite, smoke, default bench (builder ploc) both arms, ncu launch list + full capture of the three
# dominant kernels, C5 / C4 lines, kernel variants (256-bit node loads, shared-memory stack)
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -12 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
summ='
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if "unavailable" in d: print(d); continue
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f launches %s roof %s cpu %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), json.dumps(d.get("roofline"))[:700], d.get("cpu_baseline"), d.get("clocks")))
'
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1
grep -h '^{' gpurun_out/bench_ref_c3.log | python -c "$summ" "ref c3"
( time timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3_default.log 2>&1
grep -h '^{' gpurun_out/bench_c3_default.log | python -c "$summ" "c3 default"
tail -3 gpurun_out/bench_c3_default.log | grep real
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_s17_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 3 -o gpurun_out/r01_s17_full -f python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
for w in c5 c4; do for b in ploc lbvh; do
  ( time timeout 1200 python bench.py --workload $w --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_${w}_$b.log 2>&1
  grep -h '^{' gpurun_out/bench_${w}_$b.log | python -c "$summ" "$w $b"
  tail -3 gpurun_out/bench_${w}_$b.log | grep real
done; done
echo "== default lib"; timeout 600 python tools/quick_bench.py ploc ploc8 2>&1 | tee gpurun_out/quick_default.log
for v in ld256 sh8 ld256sh8; do echo "== $v"; CRT_LIB=$PWD/cudaraytracing_b200/variants/libcrt_$v.so timeout 600 python tools/quick_bench.py ploc 2>&1 | tee gpurun_out/quick_$v.log; done
ls -la gpurun_out
