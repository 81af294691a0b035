"""Development aid: frame time of the shipped configs against the k_tail switch point (CRT_TAIL)."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf

def main():
    tmp = tempfile.mkdtemp()
    for name in ("cornell-box", "veach-mis"):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
        S.set_BVH(cfg.bvh_thresh_n)
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        R = crt.Render(S, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
        for tail in sys.argv[1:] or ["0", "4096", "16384", "65536", "131072", "262144", "524288", "1048576"]:
            os.environ["CRT_TAIL"] = tail
            ms = []
            for _ in range(6):
                R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
                ms.append(R.stats()["ms_total"])
            st = R.stats()
            print("%s tail %8s: best %.3f ms median %.3f ms = %.1f Msamples/s, iterations %d launches %d" % (
                name, tail, min(ms), sorted(ms)[3], cfg.width * cfg.height * cfg.spp / min(ms) / 1e3, st["iterations"], st["kernel_launches"]), flush=True)

if __name__ == "__main__":
    main()
