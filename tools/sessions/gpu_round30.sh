#!/bin/bash
# session 30: k_shade_compat_early (queue slots reserved before the vertex arithmetic) vs k_shade<compat>: parity, A/B
mkdir -p gpurun_out
( timeout 900 python -m pytest tests -m gpu -q -x ) > gpurun_out/pytest_gpu.log 2>&1
tail -5 gpurun_out/pytest_gpu.log
( time timeout 300 python __graft_entry__.py smoke ) 2>&1 | tail -4
export QB_NO_BATCH=1
for spp in 128 16; do
  export QB_SPP=$spp QB_SCENES=cornell-box
  for e in 1 0; do echo "== early $e spp $spp"; CRT_SHADE_EARLY=$e timeout 300 python tools/quick_bench.py ploc8 2>&1 | grep "est 0" | tee gpurun_out/quick_early${e}_$spp.log; done
done
export QB_SPP=64 QB_SCENES=veach-mis
for e in 1 0; do echo "== early $e veach"; CRT_SHADE_EARLY=$e timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | grep "est 0" | tee gpurun_out/quick_early${e}_veach.log; done
for e in 1 0; do
  ( CRT_SHADE_EARLY=$e timeout 900 python bench.py --steps 3 --warmup 3 ) > gpurun_out/bench_c3_early$e.log 2>&1
  grep -h '^{' gpurun_out/bench_c3_early$e.log | python -c "
import sys, json
d = json.loads(sys.stdin.read()); print('c3 early $e', d['value'], d['roofline']['stage_ms'])"
done
