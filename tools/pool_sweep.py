"""Development aid: steady-state throughput (cornell-box 1920x1080 spp 16) against the path-pool size (CRT_POOL)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf

def main():
    tmp = tempfile.mkdtemp()
    name = "cornell-box"
    cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
    cfg = crt.load_config(cfg_path)
    d = os.path.dirname(cfg_path)
    S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
    S.set_BVH(cfg.bvh_thresh_n)
    M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 1920, 1080, 16
    for pool in sys.argv[1:] or ["262144", "524288", "1048576", "2097152", "4194304", "8388608", "16777216"]:
        os.environ["CRT_POOL"] = pool
        R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
        ms = []
        for _ in range(4):
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            ms.append(R.stats()["ms_total"])
        st = R.stats()
        R.set_stage_timing(True); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); s2 = R.stats()
        print("pool %9s: best %.3f ms = %.1f Msamples/s, iterations %d [gen %.2f ext %.2f shade %.2f shadow %.2f]" % (
            pool, min(ms), W * H * spp / min(ms) / 1e3, st["iterations"], s2["ms_generate"], s2["ms_extend"], s2["ms_shade"], s2["ms_shadow"]), flush=True)
        del R

if __name__ == "__main__":
    main()
