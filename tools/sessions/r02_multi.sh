#!/bin/bash
# round 2, multi-GPU session (gpurun --gpus N -- 'bash tools/sessions/r02_multi.sh N'): crt_group tests (buffer and PNG identical
# to one GPU), crt --gpus N, and the bench lines at N ranks (hashes must equal the N = 1 line's).
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,name --format=csv,noheader | head -8
( timeout 1200 python -m pytest tests/test_gpu_group.py -x -q 2>&1 | tail -4 )
python tools/c1_timeline.py cornell-box > /dev/null 2>&1   # warm the box
CFG=$(python -c "
import tempfile, sys
sys.path.insert(0, '.')
from tools import scene_fixture as sf
print(sf.unpack(sf.fixture('cornell-box'), tempfile.mkdtemp()))")
for g in 1 $N; do
  ./cudaraytracing_b200/crt --config $CFG --width 1920 --height 1080 --spp 64 --gpus $g --out gpurun_out/cli_g$g.png | cut -c1-400
done
cmp gpurun_out/cli_g1.png gpurun_out/cli_g$N.png && echo "crt --gpus $N: PNG identical to --gpus 1"
rm -f gpurun_out/cli_g*.png
for g in 1 $N; do
  if [ $g -eq 1 ]; then L="python bench.py"; else L="python -m torch.distributed.run --nnodes=1 --nproc-per-node $g --master-addr 127.0.0.1 --master-port 29517 bench.py"; fi
  ( time timeout 1500 $L --gpus $g --steps 3 --warmup 3 ) > gpurun_out/bench_c3_n$g.log 2>&1
  grep -h '^{' gpurun_out/bench_c3_n$g.log | python -c "
import sys, json
d = json.loads(sys.stdin.readline())
w = d.get('workloads', {})
print('N=%d C3 %.1f Msamples/s e2e %.1f frame %s accum %s' % (d['n_gpus'], d['value'], d['e2e']['value'], d['frame_sha256'][:12], d['accum_sha256'][:12]))
for k in ('c1', 'c2', 'c2_mis'):
    print('   %s %.1f (e2e %.1f) frame %s accum %s' % (k, w[k]['value'], w[k]['e2e']['value'], w[k]['frame_sha256'][:12], w[k]['accum_sha256'][:12]))
print('   c5 closest %.1f any %.1f e2e %.1f frac %.3f hits %s' % (w['c5']['value'], w['c5']['any_hit']['mrays_s'], w['c5']['e2e']['value'], w['c5']['roofline']['frac'], w['c5']['hits_sha256'][:12]))"
  grep real gpurun_out/bench_c3_n$g.log
done
