#!/bin/bash
# round 2, session 14: per-axis error terms of the folded bias - GPU suite (incl. crt_group at N = 1), C5 at 40M / 100M rays,
# steady state at 1080p / 4K, the shipped frames.
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
for n in 40000000 100000000; do
  CRT_C5_RAYS=$n timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('C5 n=$n closest', d['value'], 'any', d['any_hit']['mrays_s'], 'e2e', d['e2e']['value'], d['e2e']['pcie_frac'], d['e2e']['parity'])"
done
q() { env QB_SCENES=cornell-box QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1 | cut -c1-170; }
q QB_SPP=128
q QB_W=3840 QB_H=2160 QB_SPP=48
QB_SCENES=cornell-box QB_SPP=16 timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -1
python tools/c1_timeline.py cornell-box 2>&1 | tail -1
python tools/c1_timeline.py veach-mis 2>&1 | tail -1
