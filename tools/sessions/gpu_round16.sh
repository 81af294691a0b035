#!/bin/bash
# session 16: full GPU suite (first GPU run of the PLOC builder), smoke, default bench both arms, ncu launch list,
# builder comparison (lbvh / ploc / lbvh8 / ploc8), C1/C2 lines
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1200 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -25 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
summ='
import sys, json
for ln in sys.stdin:
    d = json.loads(ln)
    if "unavailable" in d: print(d); continue
    print(sys.argv[1], "value %.1f ms/step %.3f e2e %.1f launches %s roof %s cpu %s clocks %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d.get("gpu_launches"), json.dumps(d.get("roofline"))[:700], d.get("cpu_baseline"), d.get("clocks")))
'
echo "== quick"; timeout 900 python tools/quick_bench.py lbvh ploc lbvh8 ploc8 2>&1 | tee gpurun_out/quick_default.log
run_w() {  # workload builders...
  w=$1; shift
  ( time timeout 600 python bench.py --impl reference --workload $w --steps 3 --warmup 3 ) > gpurun_out/bench_ref_$w.log 2>&1
  grep -h '^{' gpurun_out/bench_ref_$w.log | python -c "$summ" "ref $w"
  for b in "$@"; do
    ( time timeout 900 python bench.py --workload $w --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_${w}_$b.log 2>&1
    grep -h '^{' gpurun_out/bench_${w}_$b.log | python -c "$summ" "$w $b"
    tail -3 gpurun_out/bench_${w}_$b.log | grep real
  done
}
run_w c3 lbvh ploc
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_s16_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
run_w c1 lbvh ploc
run_w c2 lbvh ploc
ls -la gpurun_out
