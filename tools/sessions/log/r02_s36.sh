#!/bin/bash
# round 2, session 36 (2 GPUs): crt_group after the warm-up at creation and the larger reduce bands; viewer and ray-batch tests; first frame of crt --gpus 2
mkdir -p gpurun_out
( timeout 900 python -m pytest tests/test_gpu_group.py tests/test_gpu_viewer.py tests/test_gpu_synthetic.py -m gpu -x -q 2>&1 | tail -5 ) | tee gpurun_out/r02_s36.log
cfg=$(python - <<'PY'
import sys, os, tempfile
sys.path.insert(0, os.getcwd())
from tools import scene_fixture as sf
print(sf.unpack(sf.fixture("cornell-box"), os.path.join(tempfile.mkdtemp(), "cornell-box")))
PY
)
for g in 1 2; do
  timeout 300 cudaraytracing_b200/crt --config $cfg --width 3840 --height 2160 --spp 64 --gpus $g --out gpurun_out/s36_g$g.png | tee -a gpurun_out/r02_s36.log
done
cmp gpurun_out/s36_g1.png gpurun_out/s36_g2.png && echo "4K spp 64: --gpus 2 PNG identical to --gpus 1" | tee -a gpurun_out/r02_s36.log
rm -f gpurun_out/s36_g*.png
( GP_W=800 GP_H=600 GP_SPP=2 timeout 600 python tools/group_probe.py; GP_SCENE=veach-mis GP_W=800 GP_H=600 GP_SPP=4 timeout 600 python tools/group_probe.py ) 2>&1 | grep "frame 3\|single frame 2" | tee -a gpurun_out/r02_s36.log
