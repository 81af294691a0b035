"""Quick end-to-end GPU check against the oracle (development aid; the formal checks are tests/)."""
import math, os, sys, tempfile, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import cudaraytracing_b200 as crt
from oracle import orc
from tools import scene_fixture as sf

def main():
    names = sys.argv[1:] or ["veach-mis", "cornell-box"]
    tmp = tempfile.mkdtemp()
    for name in names:
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        obj = os.path.join(d, cfg.OBJ_paths[0][0])
        t0 = time.time(); S = crt.Scene().add_obj(obj, d); t1 = time.time()
        ms = S.set_BVH(cfg.bvh_thresh_n); t2 = time.time()
        print(name, "load %.1f ms, build kernels %.3f ms, upload+build wall %.1f ms" % ((t1-t0)*1e3, ms, (t2-t1)*1e3), S.counts())
        O = orc.Scene().add_obj(obj, d)
        onodes, oorder, olast, obounds = O.build_new_bvh(cfg.bvh_thresh_n)
        nodes, order, last, bounds = S.export_bvh()
        print("  bvh: nodes equal", nodes.tobytes() == onodes.tobytes(), "order equal", np.array_equal(order, oorder),
              "last equal", np.array_equal(last, olast), "bounds equal", np.array_equal(bounds, obounds), len(nodes), len(onodes))
        if nodes.tobytes() != onodes.tobytes() and len(nodes) == len(onodes):
            bad = np.nonzero(nodes != onodes)[0]; print("   first diffs", bad[:5], nodes[bad[:2]], onodes[bad[:2]])
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height)
        t, face, kms = S.trace_rays(rays, crt.RAY_CLOSEST)
        ot, oface = O.trace(rays, which=0, mode=0)
        print("  primary rays: %d, kernel %.3f ms = %.1f Mrays/s; face mismatches %d, t mismatches %d, hit frac %.3f" % (
            len(rays), kms, len(rays)/kms/1e3, int((face != oface).sum()), int((t.view(np.uint32) != ot.view(np.uint32)).sum()), (face>=0).mean()))
        # random any-hit rays
        rng = np.random.default_rng(1)
        lo, hi = bounds[:3], bounds[3:]
        n = 200000
        r2 = np.zeros((n, 8), np.float32)
        r2[:, 0:3] = rng.uniform(lo, hi, (n, 3))
        dd = rng.normal(size=(n, 3)); dd /= np.linalg.norm(dd, axis=1, keepdims=True)
        r2[:, 4:7] = dd
        r2[:, 3] = rng.uniform(0, np.linalg.norm(hi-lo), n)
        for mode in (0, 1):
            t, face, kms = S.trace_rays(r2, mode)
            ot, oface = O.trace(r2, which=0, mode=mode)
            print("  random rays mode %d: kernel %.3f ms = %.1f Mrays/s; face mismatches %d, t mismatches %d" % (
                mode, kms, n/kms/1e3, int((face != oface).sum()), int((t.view(np.uint32) != ot.view(np.uint32)).sum())))
        R = crt.Render(S, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
        for rep in range(3):
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            st = R.stats()
            print("  render: %.3f ms  %.1f Msamples/s  iters %d launches %d  rays e/s/p %d/%d/%d" % (
                st["ms_total"], cfg.width*cfg.height*cfg.spp/st["ms_total"]/1e3, st["iterations"], st["kernel_launches"],
                st["extend_rays"], st["shadow_rays"], st["probe_rays"]))
        acc = R.get_accum_i64()
        t0 = time.time()
        oacc, ost = O.render(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height, 0, cfg.spp, cfg.P_RR, cfg.light_sample_n)
        print("  oracle render %.2f s; rays e/s/p %d/%d/%d" % (time.time()-t0, ost["extend_rays"], ost["shadow_rays"], ost["probe_rays"]))
        neq = int((acc != oacc).sum())
        print("  accum: %d of %d values differ; max abs diff %g (fixed-point units / 2^32)" % (neq, acc.size, float(np.abs(acc-oacc).max())/2**32))
        if neq:
            bad = np.nonzero(acc != oacc)[0][:10]; print("   first bad", bad, acc[bad], oacc[bad])
        R.set_stage_timing(True); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); st = R.stats()
        print("  stage ms: generate %.3f extend %.3f shade %.3f shadow %.3f total %.3f" % (st["ms_generate"], st["ms_extend"], st["ms_shade"], st["ms_shadow"], st["ms_total"]))
        R.set_stage_timing(False)
        os.makedirs("gpurun_out", exist_ok=True)
        R.save_frame_buffer("gpurun_out/%s.png" % name)
        if name == "cornell-box":
            for spp in (16, 64):
                R.set_spp(spp); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); st = R.stats()
                print("  spp %d: %.3f ms  %.1f Msamples/s iters %d" % (spp, st["ms_total"], cfg.width*cfg.height*spp/st["ms_total"]/1e3, st["iterations"]))
            R.save_frame_buffer("gpurun_out/%s_spp64.png" % name)

if __name__ == "__main__":
    main()
