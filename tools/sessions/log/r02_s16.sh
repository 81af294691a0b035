#!/bin/bash
# round 2, session 16: fewer IEEE divisions per vertex (both estimators) - GPU suite, steady state, shipped frames, C5 host buffers
mkdir -p gpurun_out
( timeout 1800 python -m pytest tests -m gpu -x -q 2>&1 | tail -4 )
q() { env QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -4 | cut -c1-170; }
q QB_SPP=128
q QB_W=3840 QB_H=2160 QB_SPP=48 QB_SCENES=cornell-box
python tools/c1_timeline.py cornell-box 2>&1 | tail -1
python tools/c1_timeline.py veach-mis 2>&1 | tail -1
CRT_C5_RAYS=40000000 timeout 600 python bench.py --workload c5 --steps 2 --warmup 1 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
print('C5 closest', d['value'], 'any', d['any_hit']['mrays_s'], 'e2e', d['e2e'])"
