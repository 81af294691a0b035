"""Development aid: steady-state stage times of one library build (CRT_LIB selects a variant) on both shipped scenes
at 1920x1080 spp 16, for each builder named on the command line (lbvh, lbvh8); optional C5-style batch rates."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf

def main():
    tmp = tempfile.mkdtemp()
    B = {"lbvh": 0, "lbvh8": 1, "ploc": 2, "ploc8": 3}
    builders = [a for a in sys.argv[1:] if a in B] or ["lbvh"]
    W, H, spp = int(os.environ.get("QB_W", "1920")), int(os.environ.get("QB_H", "1080")), int(os.environ.get("QB_SPP", "16"))
    for name in os.environ.get("QB_SCENES", "cornell-box,veach-mis").split(","):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        for b in builders:
            S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
            S.set_BVH(int(os.environ.get("QB_THRESH", cfg.bvh_thresh_n)), builder=B[b])
            M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
            for est in (0, 1):
                R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
                R.set_estimator(est)
                ms = []
                for _ in range(4):
                    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
                    ms.append(R.stats()["ms_total"])
                R.set_stage_timing(True); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); s2 = R.stats()
                print("%-11s %-5s est %d: best %.3f ms = %.1f Msamples/s [gen %.2f ext %.2f shade %.2f shadow %.2f] rays %d/%d/%d" % (
                    name, b, est, min(ms), W * H * spp / min(ms) / 1e3, s2["ms_generate"], s2["ms_extend"], s2["ms_shade"], s2["ms_shadow"],
                    s2["extend_rays"], s2["shadow_rays"], s2["probe_rays"]), flush=True)
                del R
            if os.environ.get("QB_NO_BATCH"):
                continue
            # incoherent batch on this scene (L2-resident BVH)
            lo, hi = S.export_bvh()[3][:3], S.export_bvh()[3][3:]
            rng = np.random.default_rng(5)
            n = 4_000_000
            r = np.zeros((n, 8), np.float32)
            r[:, 0:3] = rng.uniform(lo, hi, (n, 3))
            dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
            r[:, 4:7] = dd
            r[:, 3] = np.finfo(np.float32).max
            k0 = min(S.trace_rays(r, 0)[2] for _ in range(3))
            r[:, 3] = rng.uniform(0, np.linalg.norm(hi - lo), n)
            k1 = min(S.trace_rays(r, 1)[2] for _ in range(3))
            print("%-11s %-5s incoherent 4M rays: closest %.1f Mrays/s, any %.1f Mrays/s" % (name, b, n / k0 / 1e3, n / k1 / 1e3), flush=True)

if __name__ == "__main__":
    main()
