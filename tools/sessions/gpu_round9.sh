#!/bin/bash
mkdir -p gpurun_out
( time timeout 1500 python -m pytest tests -m gpu -x -q ) > gpurun_out/pytest_gpu.log 2>&1
tail -15 gpurun_out/pytest_gpu.log | head -12
python - <<'PY'
import os, sys, tempfile
sys.path.insert(0, os.getcwd())
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf
tmp = tempfile.mkdtemp()
for name in ("cornell-box", "veach-mis"):
    cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name)); cfg = crt.load_config(cfg_path); d = os.path.dirname(cfg_path)
    S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d); S.set_BVH(cfg.bvh_thresh_n)
    M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    for est in (0, 1):
        for (W, H, spp) in ((cfg.width, cfg.height, cfg.spp), (1920, 1080, 64)):
            R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n); R.set_estimator(est)
            ms = []
            for _ in range(4):
                R.run_view(cfg.eye_pos, M, cfg.fovy_rad); ms.append(R.stats()["ms_total"])
            print("%s est %d %dx%d spp %d: %.3f ms = %.1f Msamples/s" % (name, est, W, H, spp, min(ms), W * H * spp / min(ms) / 1e3), flush=True)
            if spp == 64: R.save_frame_buffer("gpurun_out/%s_est%d.png" % (name, est))
PY
