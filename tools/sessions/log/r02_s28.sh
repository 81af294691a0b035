#!/bin/bash
# round 2, session 28 (2 GPUs): crt_group with the reduce in row bands + overlapped resolve; the viewer shim tests
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_viewer.py -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/r02_s28.log
cfg=$(python - <<'PY'
import sys, os, tempfile
sys.path.insert(0, os.getcwd())
from tools import scene_fixture as sf
print(sf.unpack(sf.fixture("cornell-box"), os.path.join(tempfile.mkdtemp(), "cornell-box")))
PY
)
for g in 1 2; do
  timeout 300 cudaraytracing_b200/crt --config $cfg --width 3840 --height 2160 --spp 64 --gpus $g --out gpurun_out/s28_g$g.png | tee -a gpurun_out/r02_s28.log
done
cmp gpurun_out/s28_g1.png gpurun_out/s28_g2.png && echo "4K spp 64: --gpus 2 PNG identical to --gpus 1" | tee -a gpurun_out/r02_s28.log
rm -f gpurun_out/s28_g*.png
