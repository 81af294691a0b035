"""Development aid: frame times of crt_group (one host thread, N GPUs) against the single-GPU render object, frame by frame
(the first frame of a process pays allocations and module loading)."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf

def main():
    name = os.environ.get("GP_SCENE", "cornell-box")
    W, H, spp = int(os.environ.get("GP_W", "3840")), int(os.environ.get("GP_H", "2160")), int(os.environ.get("GP_SPP", "64"))
    cfg_path = sf.unpack(sf.fixture(name), os.path.join(tempfile.mkdtemp(), name))
    cfg = crt.load_config(cfg_path)
    d = os.path.dirname(cfg_path)
    S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
    S.set_BVH(cfg.bvh_thresh_n, builder=3, device=0)
    M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    for f in range(3):
        t0 = time.time(); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); t1 = time.time()
        print("%s %dx%d spp %d single frame %d: wall %.2f ms, gpu %.2f ms" % (name, W, H, spp, f, (t1 - t0) * 1e3, R.stats()["ms_total"]), flush=True)
    del R
    for n in [int(x) for x in os.environ.get("GP_N", "1,2").split(",")]:
        if n > crt.device_count():
            continue
        G = crt.RenderGroup(S, W, H, list(range(n)), spp, cfg.P_RR, cfg.light_sample_n)
        for f in range(4):
            t0 = time.time(); G.run_view(cfg.eye_pos, M, cfg.fovy_rad); t1 = time.time()
            per = [G.stats(k) for k in range(n)]
            print("group N=%d frame %d: wall %.2f ms, gpu ms %s, reduce %.3f ms, iterations %s" % (
                n, f, (t1 - t0) * 1e3, ["%.1f" % s["ms_total"] for s in per], per[0].get("reduce_ms", -1), [s["iterations"] for s in per]), flush=True)
        G.close()

if __name__ == "__main__":
    main()
