#!/bin/bash
# round 2, session 8: real schedule of a C1 / C2 frame (event per launch), tile size at 4K, GPU suite of the simplified fetch.
mkdir -p gpurun_out
q() { env QB_SCENES=cornell-box QB_NO_BATCH=1 "$@" timeout 600 python tools/quick_bench.py ploc8 2>&1 | tail -2 | head -1 | cut -c1-170; }
echo "== C1 timeline"; CRT_TIMELINE=1 python tools/c1_timeline.py cornell-box 2>&1 | tail -40
echo "== C1 no timeline"; python tools/c1_timeline.py cornell-box 2>&1 | tail -4
echo "== C2 timeline"; CRT_TIMELINE=1 python tools/c1_timeline.py veach-mis 2>&1 | tail -36
echo "== 4K tile 2^21 / 2^20 / 2^19 / 2^22"
q QB_W=3840 QB_H=2160 QB_SPP=48
q QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=1048576
q QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=524288
q QB_W=3840 QB_H=2160 QB_SPP=48 CRT_TILE_PX=4194304
echo "== 1080p tile off / 2^20"
q QB_SPP=128
q QB_SPP=128 CRT_TILE_PX=1048576
( timeout 1500 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 )
