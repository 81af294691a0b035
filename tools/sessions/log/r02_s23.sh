#!/bin/bash
# round 2, session 23: ncu of the tail path tracer on C1 (six launches: the long one is the real hand-over), GPU suite, both bench arms
mkdir -p gpurun_out
timeout 600 ncu --set full --metrics lts__t_bytes.sum --clock-control none --import-source on -k regex:'k_tail' -s 3 -c 6 -o gpurun_out/r02_final_c1 -f python tools/c1_timeline.py cornell-box > gpurun_out/ncu_c1.log 2>&1
tail -1 gpurun_out/ncu_c1.log | cut -c1-120
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1; grep real gpurun_out/bench_ref_c3.log
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3.log 2>&1; grep real gpurun_out/bench_c3.log; tail -2 gpurun_out/bench_c3.log | cut -c1-300
grep -h '^{' gpurun_out/bench_c3.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
w = d.get('workloads', {})
print('C3 %.1f Msamples/s e2e %.1f  roofline %s' % (d['value'], d['e2e']['value'], {k: d['roofline'][k] for k in ('bound', 'frac', 'fp32_frac', 'issue_frac', 'l2_gbs', 'lane_efficiency', 'traffic')}))
for k in ('c1', 'c2', 'c2_mis', 'c4'):
    print('   %s %.1f (e2e %.1f) %.3f ms' % (k, w[k]['value'], w[k]['e2e']['value'], w[k]['ms_per_step']))
print('   c5 closest %.1f any %.1f e2e %.1f' % (w['c5']['value'], w['c5']['any_hit']['mrays_s'], w['c5']['e2e']['value']))"
