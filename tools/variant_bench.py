"""Times kernel variants (cudaraytracing_b200/variants/*.so, selected with CRT_LIB) on the two scenes."""
import glob, os, subprocess, sys, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

def child():
    import cudaraytracing_b200 as crt
    from tools import scene_fixture as sf
    tmp = tempfile.mkdtemp()
    out = []
    for name, spp in (("cornell-box", 16), ("veach-mis", 16), ("cornell-box", 2)):
        cfg_path = sf.unpack(sf.fixture(name), os.path.join(tmp, name + str(spp)))
        cfg = crt.load_config(cfg_path)
        d = os.path.dirname(cfg_path)
        S = crt.Scene().add_obj(os.path.join(d, cfg.OBJ_paths[0][0]), d)
        S.set_BVH(cfg.bvh_thresh_n)
        M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
        R = crt.Render(S, cfg.width, cfg.height, spp, cfg.P_RR, cfg.light_sample_n)
        best = 1e9
        for _ in range(4):
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            best = min(best, R.stats()["ms_total"])
        R.set_stage_timing(True); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); st = R.stats()
        out.append("%s spp%d: %.3f ms (%.0f Msamples/s) [ext %.2f shade %.2f shadow %.2f]" % (
            name[:6], spp, best, cfg.width * cfg.height * spp / best / 1e3, st["ms_extend"], st["ms_shade"], st["ms_shadow"]))
    print(" | ".join(out))

if __name__ == "__main__":
    if len(sys.argv) > 1 and sys.argv[1] == "child":
        child()
    else:
        for so in sorted(glob.glob(os.path.join(ROOT, "cudaraytracing_b200", "variants", "*.so"))):
            env = dict(os.environ, CRT_LIB=so)
            r = subprocess.run([sys.executable, __file__, "child"], env=env, capture_output=True, text=True)
            print(os.path.basename(so), r.stdout.strip() or r.stderr[-300:], flush=True)
