"""Development aid: parity + timing of a traversal variant (CRT_LIB) against the oracle."""
import os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from oracle import orc
from tools import scene_fixture as sf

def main():
    tmp = tempfile.mkdtemp()
    for name in ("veach-mis", "cornell-box"):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        obj = os.path.join(d, cfg.OBJ_paths[0][0])
        S = crt.Scene().add_obj(obj, d); S.set_BVH(cfg.bvh_thresh_n)
        O = orc.Scene().add_obj(obj, d); O.build_new_bvh(cfg.bvh_thresh_n)
        _, _, _, bounds = S.export_bvh()
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height)
        t, f, kms = S.trace_rays(rays, 0); ot, of = O.trace(rays, which=0, mode=0)
        print(name, "primary: face mismatches %d t mismatches %d (%.1f Mrays/s)" % ((f != of).sum(), (t.view(np.uint32) != ot.view(np.uint32)).sum(), len(rays) / kms / 1e3))
        rng = np.random.default_rng(1); n = 300000
        lo, hi = bounds[:3], bounds[3:]
        r2 = np.zeros((n, 8), np.float32); r2[:, 0:3] = rng.uniform(lo, hi, (n, 3))
        dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True); r2[:, 4:7] = dd
        r2[:, 3] = np.float32(3.4e38)
        t, f, kms = S.trace_rays(r2, 0); ot, of = O.trace(r2, which=0, mode=0)
        print("  random closest: face mismatches %d t mismatches %d (%.1f Mrays/s)" % ((f != of).sum(), (t.view(np.uint32) != ot.view(np.uint32)).sum(), n / kms / 1e3))
        r2[:, 3] = rng.uniform(0, np.linalg.norm(hi - lo), n)
        t, f, kms = S.trace_rays(r2, 1); ot, of = O.trace(r2, which=0, mode=1)
        print("  random any: blocked-status mismatches %d, face differs %d (%.1f Mrays/s)" % (((f >= 0) != (of >= 0)).sum(), (f != of).sum(), n / kms / 1e3))
        R = crt.Render(S, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
        R.run_view(cfg.eye_pos, M, cfg.fovy_rad); acc = R.get_accum_i64(); st = R.stats()
        oacc, ost = O.render(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height, 0, cfg.spp, cfg.P_RR, cfg.light_sample_n)
        print("  render: accum values differing %d; rays equal %s; %.3f ms" % ((acc != oacc).sum(),
              (st["extend_rays"], st["shadow_rays"], st["probe_rays"]) == (ost["extend_rays"], ost["shadow_rays"], ost["probe_rays"]), st["ms_total"]), flush=True)

if __name__ == "__main__":
    main()
