#!/bin/bash
# round 2, session 41: the final binary (gate in the any-hit batch kernel only) - GPU suite, default bench line
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3.log 2>&1; grep real gpurun_out/bench_c3.log
grep -h '^{' gpurun_out/bench_c3.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
w = d.get('workloads', {})
print('C3 %.1f Msamples/s e2e %.1f  roofline %s' % (d['value'], d['e2e']['value'], {k: d['roofline'].get(k) for k in ('bound', 'frac', 'fp32_frac', 'issue_frac', 'l2_gbs', 'lane_efficiency', 'traffic')}))
for k in ('c1', 'c2', 'c2_mis', 'c4'):
    print('   %s %.1f (e2e %.1f) %.3f ms' % (k, w[k]['value'], w[k]['e2e']['value'], w[k]['ms_per_step']))
print('   c5 closest %.1f any %.1f e2e %.1f' % (w['c5']['value'], w['c5']['any_hit']['mrays_s'], w['c5']['e2e']['value']))"
python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
