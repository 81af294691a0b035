#!/bin/bash
# round 2, session 19: tail launched first in its iteration (beside k_shadow of the previous one); default bench line with the C4 sub-measurement
mkdir -p gpurun_out
for i in 1 2; do python tools/c1_timeline.py cornell-box 2>&1 | tail -1; python tools/c1_timeline.py veach-mis 2>&1 | tail -1; done
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3.log 2>&1; grep real gpurun_out/bench_c3.log; tail -2 gpurun_out/bench_c3.log | cut -c1-300
grep -h '^{' gpurun_out/bench_c3.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
w = d.get('workloads', {})
print('C3 %.1f Msamples/s e2e %.1f' % (d['value'], d['e2e']['value']))
for k in ('c1', 'c2', 'c2_mis', 'c4'):
    print('   %s %.1f (e2e %.1f) %.3f ms' % (k, w[k]['value'], w[k]['e2e']['value'], w[k]['ms_per_step']))
print('   c5 closest %.1f any %.1f e2e %.1f' % (w['c5']['value'], w['c5']['any_hit']['mrays_s'], w['c5']['e2e']['value']))"
