"""Development aid: the real schedule of one frame of a shipped config (CRT_TIMELINE=1 prints an event per launch)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf
name = sys.argv[1] if len(sys.argv) > 1 else "cornell-box"
tmp = tempfile.mkdtemp()
cfg = crt.load_config(sf.unpack(sf.fixture(name), tmp))
S = crt.Scene().add_obj(os.path.join(tmp, cfg.OBJ_paths[0][0]), tmp)
S.set_BVH(cfg.bvh_thresh_n, builder=3)
M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
R = crt.Render(S, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
for k in range(4):
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    st = R.stats()
    print("frame %d: %.3f ms, %d iterations, %d launches, rays %d/%d" % (k, st["ms_total"], st["iterations"], st["kernel_launches"], st["extend_rays"], st["shadow_rays"]), file=sys.stderr)
