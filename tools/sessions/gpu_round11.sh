#!/bin/bash
# session 11: full GPU test suite (incl. the wide-BVH tests), default bench both arms, ncu launch list + full capture
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv,noheader
( time timeout 1700 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -6 gpurun_out/pytest_gpu.log
( time timeout 600 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1
grep -h '^{' gpurun_out/bench_ref_c3.log | cut -c1-400
for b in lbvh lbvh8; do
  ( time timeout 900 python bench.py --steps 3 --warmup 3 --builder $b ) > gpurun_out/bench_c3_$b.log 2>&1
  grep -h '^{' gpurun_out/bench_c3_$b.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('c3 $b value %.1f ms/step %.3f e2e %.1f launches %d roof %s cpu %s clocks %s' % (d['value'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches'], json.dumps(d['roofline'])[:700], d['cpu_baseline'], d['clocks']))
"
done
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r01_s11_launches.csv python bench.py --steps 1 --warmup 1 > gpurun_out/ncu_launch.log 2>&1
tail -2 gpurun_out/ncu_launch.log | cut -c1-300
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'k_extend|k_shadow|k_shade' -s 60 -c 6 -o gpurun_out/r01_s11_full -f python bench.py --steps 1 --warmup 0 > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-300
ls -la gpurun_out
