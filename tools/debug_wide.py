"""Development aid (GPU box): where does the wide-BVH render differ from the pair-BVH render / the oracle?"""
import os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudaraytracing_b200 as crt
from oracle import orc
from tools import scene_fixture as sf

def main():
    tmp = tempfile.mkdtemp()
    out = {}
    for name in ("veach-mis", "cornell-box"):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path); d = os.path.dirname(cfg_path)
        obj = os.path.join(d, cfg.OBJ_paths[0][0])
        P = crt.Scene().add_obj(obj, d); P.set_BVH(cfg.bvh_thresh_n, builder=0)
        Wd = crt.Scene().add_obj(obj, d); Wd.set_BVH(cfg.bvh_thresh_n, builder=1)
        O = orc.Scene().add_obj(obj, d); O.build_new_bvh(cfg.bvh_thresh_n)
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        for est in (0, 1):
            for tail in ("131072", "0", "100000000"):
                os.environ["CRT_TAIL"] = tail
                acc = []
                for S in (P, Wd):
                    R = crt.Render(S, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n); R.set_estimator(est)
                    R.run_view(cfg.eye_pos, M, cfg.fovy_rad); acc.append(R.get_accum_i64().reshape(-1, 3)); del R
                bad = np.nonzero((acc[0] != acc[1]).any(axis=1))[0]
                print("%s est %d tail %s: pixels differing wide vs pair: %d" % (name, est, tail, len(bad)), flush=True)
                for p in bad[:6]:
                    print("    pixel %d (x %d y %d): pair %s wide %s diff %s" % (p, p % cfg.width, p // cfg.width, acc[0][p], acc[1][p], acc[1][p] - acc[0][p]))
        os.environ.pop("CRT_TAIL")
        # secondary rays: origins on surfaces (primary hit points), random directions / directions toward light triangles
        rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height)
        t, f, _ = P.trace_rays(rays, 0)
        hit = f >= 0
        pos = rays[hit, 0:3] + t[hit, None] * rays[hit, 4:7]
        rng = np.random.default_rng(7)
        tv = O.tris()["verts"].reshape(-1, 3, 3)
        lfaces = np.concatenate([fa for fa, _ in O.lights()])
        for rep in range(3):       # shadow-like rays: surface point -> point on a light triangle, tmax = distance
            n = len(pos)
            k = lfaces[rng.integers(0, len(lfaces), n)]
            a = rng.uniform(0, 1, (n, 1)).astype(np.float32); b = (rng.uniform(0, 1, (n, 1)) * (1 - a)).astype(np.float32)
            lp = a * tv[k, 0] + b * tv[k, 1] + (1 - a - b) * tv[k, 2]
            dist = (lp - pos).astype(np.float32)
            ln = np.sqrt((dist * dist).sum(axis=1, keepdims=True)).astype(np.float32)
            r = np.zeros((n, 8), np.float32)
            r[:, 0:3] = pos; r[:, 4:7] = dist / ln
            nrm = np.sqrt((r[:, 4:7] ** 2).sum(axis=1, keepdims=True)).astype(np.float32)
            r[:, 4:7] /= nrm
            with np.errstate(all="ignore"):
                r[:, 3] = dist[:, 0] / r[:, 4]
            ok = np.isfinite(r).all(axis=1)
            r = r[ok]
            tp, fp, _ = P.trace_rays(r, 1); tw, fw, _ = Wd.trace_rays(r, 1)
            ot, of = O.trace(r, which=0, mode=1)
            bw = np.nonzero((fw >= 0) != (of >= 0))[0]; bp = np.nonzero((fp >= 0) != (of >= 0))[0]
            print("%s shadow-like rep %d: %d rays, wide mismatches %d, pair mismatches %d" % (name, rep, len(r), len(bw), len(bp)), flush=True)
            if len(bw):
                out["%s_shadow_%d" % (name, rep)] = r[bw[:64]]
                for kk in bw[:4]:
                    print("    ray", r[kk], "wide", tw[kk], fw[kk], "oracle", ot[kk], of[kk])
        for rep in range(6):
            n = len(pos)
            r = np.zeros((n, 8), np.float32)
            r[:, 0:3] = pos
            dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
            if rep >= 3:   # axis-parallel / planar directions
                dd[:, rep - 3] = 0.0; dd /= np.linalg.norm(dd, axis=1, keepdims=True)
            r[:, 4:7] = dd.astype(np.float32)
            nrm = np.linalg.norm(r[:, 4:7].astype(np.float32), axis=1, keepdims=True).astype(np.float32)
            r[:, 4:7] /= nrm
            r[:, 3] = np.finfo(np.float32).max
            tp, fp, _ = P.trace_rays(r, 0); tw, fw, _ = Wd.trace_rays(r, 0)
            ot, of = O.trace(r, which=0, mode=0)
            bw = np.nonzero((fw != of) | (tw.view(np.uint32) != ot.view(np.uint32)))[0]
            bp = np.nonzero((fp != of) | (tp.view(np.uint32) != ot.view(np.uint32)))[0]
            print("%s rep %d closest: %d rays, wide mismatches %d, pair mismatches %d" % (name, rep, n, len(bw), len(bp)), flush=True)
            if len(bw):
                out["%s_closest_%d" % (name, rep)] = r[bw[:64]]
                for k in bw[:4]:
                    print("    ray", r[k], "wide", tw[k], fw[k], "oracle", ot[k], of[k])
            r[:, 3] = rng.uniform(0, 30.0 if name == "veach-mis" else 900.0, n).astype(np.float32)
            tp, fp, _ = P.trace_rays(r, 1); tw, fw, _ = Wd.trace_rays(r, 1)
            ot, of = O.trace(r, which=0, mode=1)
            bw = np.nonzero((fw >= 0) != (of >= 0))[0]; bp = np.nonzero((fp >= 0) != (of >= 0))[0]
            print("%s rep %d any: wide mismatches %d, pair mismatches %d" % (name, rep, len(bw), len(bp)), flush=True)
            if len(bw):
                out["%s_any_%d" % (name, rep)] = r[bw[:64]]
                for k in bw[:4]:
                    print("    ray", r[k], "wide", tw[k], fw[k], "oracle", ot[k], of[k])
    if out:
        np.savez("gpurun_out/wide_mismatch.npz", **out)

if __name__ == "__main__":
    main()
