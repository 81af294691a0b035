"""Converged-image study: `compat` against the reference's own view_render_kernel at 240x180 for a ladder
of spp. Per spp: per-pixel RMSE and RMSE after an 8x8 box filter of (ours - reference) on the tone-mapped RGB8 frames, and the
same between two of OUR renders with different seeds (pure Monte-Carlo noise at that spp: what is left above it is bias between
the two estimators). At the top spp: the signed 8x8-box residual map (PNG, 128 = 0, 16 code values per unit) and its mean over
the pixels whose primary hit is directly lit / not lit (the reference's rounding-dependent self-occlusion of light samples, E8,
can only show where a light sample would otherwise arrive).
usage: converged_study.py [spp ...]   -> gpurun_out/r02_converged.md, gpurun_out/r02_residual_<scene>.png"""
import math, os, sys, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
import cudaraytracing_b200 as crt
from oracle import orc, ref
from tools import scene_fixture as sf

W, H = 240, 180
ladder = [int(x) for x in sys.argv[1:]] or [256, 1024, 4096, 16384]


def box(img, k=8):
    h, w, c = img.shape
    return img[:h - h % k, :w - w % k].reshape(h // k, k, w // k, k, c).astype(np.float64).mean(axis=(1, 3))


def rm(a, b):
    return math.sqrt(((a.astype(np.float64) - b.astype(np.float64)) ** 2).mean())


lines = ["# r02 — converged `compat` image against the reference's kernel (240x180, tone-mapped RGB8, code values)\n",
         "`tools/converged_study.py` on one B200; reference = `view_render_kernel` from oracle/_ref/libref.so (clock()-seeded).\n"]
tmp = tempfile.mkdtemp()
for name in ("cornell-box", "veach-mis"):
    d = os.path.join(tmp, name)
    cfg = crt.load_config(sf.unpack(sf.fixture(name), d))
    obj = os.path.join(d, cfg.OBJ_paths[0][0])
    rs = ref.RefScene(obj, d, W, H, cfg.bvh_thresh_n)
    assert rs.device_init() == 0
    Mr = rs.inverse_view(cfg.eye_pos, cfg.lookat, cfg.up)
    S = crt.Scene().add_obj(obj, d)
    S.set_BVH(cfg.bvh_thresh_n, builder=3)
    M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    lines += ["\n## %s\n" % name, "| spp | RMSE ours-ref | 8x8 box ours-ref | 8x8 box ours-ours (two seeds: noise) | 8x8 box ref-ref (two runs) | mean ratio |",
              "|---:|---:|---:|---:|---:|---:|"]
    last = None
    for spp in ladder:
        ref_img, _, _ = rs.render(cfg.eye_pos, Mr, cfg.fovy_rad, spp, cfg.P_RR, cfg.light_sample_n)
        ref_img2, _, _ = rs.render(cfg.eye_pos, Mr, cfg.fovy_rad, spp, cfg.P_RR, cfg.light_sample_n)
        imgs = []
        for seed in (0, 1):
            R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
            R.set_seed(seed)
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            imgs.append(R.get_frame_buffer())
            del R
        ratio = imgs[0].astype(np.float64).mean() / ref_img.astype(np.float64).mean()
        lines.append("| %d | %.3f | %.3f | %.3f | %.3f | %.4f |" % (spp, rm(imgs[0], ref_img), rm(box(imgs[0]), box(ref_img)), rm(box(imgs[0]), box(imgs[1])),
                                                                    rm(box(ref_img), box(ref_img2)), ratio))
        print(lines[-1], flush=True)
        last = (imgs[0], ref_img)
    rs.close()
    # signed residual at the top spp, and where it sits: directly lit primary hits against the rest
    res = box(last[0]) - box(last[1])
    png = np.clip(128.0 + 16.0 * res, 0, 255).astype(np.uint8)
    png = np.repeat(np.repeat(png, 8, axis=0), 8, axis=1)
    os.makedirs("gpurun_out", exist_ok=True)
    crt.write_png("gpurun_out/r02_residual_%s.png" % name, np.ascontiguousarray(png))
    O = orc.Scene().add_obj(obj, d)
    O.build_new_bvh(cfg.bvh_thresh_n)
    rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), W, H)
    t, face = O.trace(rays, which=0)
    tri = O.tris()
    lights = O.lights()
    lit = np.zeros(W * H, bool)
    hitp = rays[:, 0:3] + t[:, None] * rays[:, 4:7]
    for faces, _ in lights:
        cen = tri["verts"][faces].reshape(-1, 3, 3).mean(axis=(0, 1))
        dvec = cen[None, :] - hitp
        dist = np.linalg.norm(dvec, axis=1)
        sh = np.zeros((W * H, 8), np.float32)
        sh[:, 0:3] = hitp + 1e-3 * dvec / np.maximum(dist, 1e-9)[:, None]
        sh[:, 3] = dist * 0.98
        sh[:, 4:7] = dvec / np.maximum(dist, 1e-9)[:, None]
        _, bf = O.trace(sh.astype(np.float32), which=0, mode=1)
        nrm = tri["normal"][np.maximum(face, 0)]
        facing = (nrm * sh[:, 4:7]).sum(axis=1) > 0
        lit |= (face >= 0) & (bf < 0) & facing
    litb = box(lit.reshape(H, W, 1).astype(np.float64))[:, :, 0] > 0.5
    lines.append("\nAt %d spp: mean signed 8x8 residual (ours - reference) over directly lit blocks %.3f (n = %d), over the other blocks %.3f (n = %d); "
                 "RMS of the residual over lit blocks %.3f, over the others %.3f. Map: `r02_residual_%s.png` (128 = 0, 16 levels per code value)."
                 % (ladder[-1], res[litb].mean(), int(litb.sum()), res[~litb].mean(), int((~litb).sum()), math.sqrt((res[litb] ** 2).mean()),
                    math.sqrt((res[~litb] ** 2).mean()), name))
    print(lines[-1], flush=True)
open("gpurun_out/r02_converged.md", "w").write("\n".join(lines) + "\n")
