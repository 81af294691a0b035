#!/bin/bash
mkdir -p gpurun_out
for N in 1 2; do
  if [ $N = 1 ]; then L="python"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"; fi
  for w in c3 c1 c2; do
    ( time timeout 900 $L bench.py --gpus $N --workload $w --steps 3 --warmup 3 ) > gpurun_out/scale_${w}_n$N.log 2>&1
    grep -h '^{' gpurun_out/scale_${w}_n$N.log | python -c "
import sys, json
for ln in sys.stdin:
    d = json.loads(ln); print('$w N=$N value %.1f %s ms/step %.3f e2e %.1f launches %d' % (d['value'], d['unit'], d['ms_per_step'], d['e2e']['value'], d['gpu_launches']))
"
  done
done
