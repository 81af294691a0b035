#!/bin/bash
# session 24: k_shadow(k) on a second stream beside prepare/generate/extend(k+1): A/B (CRT_OVERLAP=0/1), full GPU suite
mkdir -p gpurun_out
export QB_NO_BATCH=1
for spp in 128 16; do
  export QB_SPP=$spp QB_SCENES=cornell-box
  for ov in 1 0; do echo "== overlap $ov spp $spp"; CRT_OVERLAP=$ov timeout 300 python tools/quick_bench.py ploc8 ploc 2>&1 | tee gpurun_out/quick_ov${ov}_$spp.log; done
done
export QB_SPP=64 QB_SCENES=veach-mis
for ov in 1 0; do echo "== overlap $ov veach"; CRT_OVERLAP=$ov timeout 300 python tools/quick_bench.py ploc8 2>&1 | tee gpurun_out/quick_ov${ov}_veach.log; done
( timeout 900 python -m pytest tests -m gpu -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -3 gpurun_out/pytest_gpu.log
for w in c1 c2; do for ov in 1 0; do
  ( CRT_OVERLAP=$ov timeout 600 python bench.py --workload $w --steps 5 --warmup 3 ) > gpurun_out/bench_${w}_ov$ov.log 2>&1
  grep -h '^{' gpurun_out/bench_${w}_ov$ov.log | cut -c1-120
done; done
