"""Small invocations of every kernel family for compute-sanitizer (tools/sessions/r02_sanitizer.sh): BVH build with the
four builders, a compat and a mis render (wavefront iterations, the tail path tracer, SPECULAR probes on veach-mis),
device-resident and host-buffer ray batches in both modes. Sizes are tiny: the tools slow kernels down 10-100x."""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudaraytracing_b200 as crt
from tools import scene_fixture as sf

W, H = int(os.environ.get("SAN_W", "64")), int(os.environ.get("SAN_H", "48"))
builders = [int(b) for b in os.environ.get("SAN_BUILDERS", "0,1,2,3").split(",")]
tmp = tempfile.mkdtemp()
for name in os.environ.get("SAN_SCENES", "veach-mis").split(","):
    cfg = crt.load_config(sf.unpack(sf.fixture(name), os.path.join(tmp, name)))
    d = os.path.join(tmp, name)
    for b in builders:
        S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
        S.set_BVH(cfg.bvh_thresh_n, builder=b)
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        for est in (0, 1):
            R = crt.Render(S, W, H, 2, cfg.P_RR, cfg.light_sample_n)
            R.set_estimator(est)
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            st = R.stats()
            print("%s builder %d est %d: %d iterations, rays %d/%d/%d" % (name, b, est, st["iterations"], st["extend_rays"], st["shadow_rays"], st["probe_rays"]), flush=True)
            R.get_frame_buffer()
            del R
        bounds = S.export_bvh()[3]
        rng = np.random.default_rng(3)
        n = 3000
        r = np.zeros((n, 8), np.float32)
        r[:, 0:3] = rng.uniform(bounds[:3], bounds[3:], (n, 3))
        dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        r[:, 4:7] = dd
        r[:, 3] = np.finfo(np.float32).max
        t, f, _ = S.trace_rays(r, crt.RAY_CLOSEST)
        r[:, 3] = rng.uniform(0, 10, n)
        t2, f2, _ = S.trace_rays(r, crt.RAY_ANY)
        print("%s builder %d batches: %d closest hits, %d blocked" % (name, b, int((f >= 0).sum()), int((f2 >= 0).sum())), flush=True)
        del S
print("done")
