#!/bin/bash
# round 2, session 34: final binary - launch list of the default bench command, GPU suite, both bench arms
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file gpurun_out/r02_final_launches.csv python bench.py --steps 1 --warmup 1 --no-sub > gpurun_out/ncu_launch.log 2>&1
tail -1 gpurun_out/ncu_launch.log | cut -c1-120
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
( time timeout 1200 python bench.py --impl reference --steps 3 --warmup 3 ) > gpurun_out/bench_ref_c3.log 2>&1; grep real gpurun_out/bench_ref_c3.log
( time timeout 1200 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3.log 2>&1; grep real gpurun_out/bench_c3.log
grep -h '^{' gpurun_out/bench_c3.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
w = d.get('workloads', {})
print('C3 %.1f Msamples/s e2e %.1f  roofline %s' % (d['value'], d['e2e']['value'], {k: d['roofline'].get(k) for k in ('bound', 'frac', 'fp32_frac', 'issue_frac', 'l2_gbs', 'lane_efficiency', 'traffic')}))
for k in ('c1', 'c2', 'c2_mis', 'c4'):
    print('   %s %.1f (e2e %.1f) %.3f ms' % (k, w[k]['value'], w[k]['e2e']['value'], w[k]['ms_per_step']))
print('   c5 closest %.1f any %.1f e2e %.1f pageable %.1f pcie_frac %.3f' % (w['c5']['value'], w['c5']['any_hit']['mrays_s'], w['c5']['e2e']['value'], w['c5']['e2e']['pageable_mrays_s'], w['c5']['e2e']['pcie_frac']))"
