"""GPU tests against the REFERENCE's own CUDA code (oracle/_ref/libref.so = its headers rebuilt headless for
sm_100a; built in the container that has /root/reference, shipped to the GPU box with the snapshot)."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def ref_mod(crt):
    from oracle import ref
    if crt.device_count() < 1:
        pytest.fail("no CUDA device")
    if not ref.available():
        pytest.skip("oracle/_ref/libref.so was not built (no /root/reference at build time)")
    return ref


def _box(img, k):
    h, w, c = img.shape
    return img[:h - h % k, :w - w % k].reshape(h // k, k, w // k, k, c).astype(np.float64).mean(axis=(1, 3))


@pytest.mark.parametrize("name,rmse_max,rmse_box_max,ratio_tol", [("cornell-box", 8.5, 2.8, 0.015), ("veach-mis", 2.6, 0.45, 0.005)])
def test_converged_image_vs_reference_kernel(crt, ref_mod, scene_files, name, rmse_max, rmse_box_max, ratio_tol):
    """North-star check 3: the converged `compat` image against a high-spp render of the reference's own
    view_render_kernel. Both are Monte-Carlo estimates with independent noise (the reference seeds with
    clock()), compared on the tone-mapped RGB8 frame, the only thing the reference exposes (Render.cuh:350).
    Thresholds in 8-bit code values at 2048 spp, from the spp ladder in profiles/r02_converged.md (256 ... 16384 spp,
    both sides, with the noise floor from two of our own renders and two of the reference's):
      veach-mis   per-pixel RMSE 2.85 (1024 spp) -> 1.44 (4096), 8x8-box RMSE 0.35 -> 0.19: pure noise, it keeps falling
                  with spp down to 0.11 at 16384 where two of our own renders differ by 0.10; mean ratio 1.0005-1.0009;
      cornell-box per-pixel RMSE 9.3 -> 5.9, 8x8-box RMSE 2.50 -> 2.39 -> 2.34: a bias floor of ~2.3 code values that does
                  not fall with spp (noise floor 0.41 at 16384), mean ratio 1.004. It sits on the directly lit surfaces
                  (RMS of the box residual 3.26 there, 1.13 elsewhere): the reference's shadow test compares
                  t_to_light - hit.t with an absolute 1e-5 at t ~ 400 (Render.cuh:19-27,272), so the light's own
                  triangles occlude a rounding-dependent share of its samples - 26 % on average, 12-56 % depending on
                  where the shaded point is, with this repo's arithmetic - and nvcc's FMA contraction of the reference
                  shifts that pattern by a few per cent of the direct light."""
    f = scene_files[name]
    cfg = crt.load_config(f["cfg_path"])
    W, H, spp = 240, 180, 2048
    rs = ref_mod.RefScene(f["obj"], f["dir"], W, H, cfg.bvh_thresh_n)
    assert rs.device_init() == 0
    M = rs.inverse_view(cfg.eye_pos, cfg.lookat, cfg.up)
    ref_img, kms, _ = rs.render(cfg.eye_pos, M, cfg.fovy_rad, spp, cfg.P_RR, cfg.light_sample_n)
    rs.close()
    S = crt.Scene().add_obj(f["obj"], f["dir"])
    S.set_BVH(cfg.bvh_thresh_n)
    R = crt.Render(S, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up), cfg.fovy_rad)
    img = R.get_frame_buffer()
    d = img.astype(np.float64) - ref_img.astype(np.float64)
    rmse = math.sqrt((d ** 2).mean())
    rmse_box = math.sqrt(((_box(img, 8) - _box(ref_img, 8)) ** 2).mean())
    ratio = img.astype(np.float64).mean() / ref_img.astype(np.float64).mean()
    print("%s: rmse %.3f  rmse(8x8 box) %.3f  mean ratio %.4f  ref kernel %.1f ms ours %.1f ms" % (name, rmse, rmse_box, ratio, kms, R.stats()["ms_total"]))
    import os
    os.makedirs("gpurun_out", exist_ok=True)
    crt.write_png("gpurun_out/conv_%s_ours.png" % name, img)
    crt.write_png("gpurun_out/conv_%s_ref.png" % name, ref_img)
    assert rmse <= rmse_max and rmse_box <= rmse_box_max and abs(ratio - 1.0) <= ratio_tol


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_hit_distance_vs_reference_device_traversal(crt, ref_mod, orc, scene_files, name):
    """Primary-ray hit distances from the reference's own DeviceBVH::intersect running on the GPU (its BVH, its
    traversal, nvcc's FMA contraction) against ours. Arithmetic differs (contraction), so t agrees to a few ulp,
    and hit/miss agrees except for rays within rounding of an edge."""
    f = scene_files[name]
    cfg = crt.load_config(f["cfg_path"])
    W, H = 400, 300
    rs = ref_mod.RefScene(f["obj"], f["dir"], W, H, cfg.bvh_thresh_n)
    assert rs.device_init() == 0
    M = crt.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), W, H)
    t_ref, _ = rs.trace(rays)
    rs.close()
    S = crt.Scene().add_obj(f["obj"], f["dir"])
    S.set_BVH(cfg.bvh_thresh_n)
    t, face, _ = S.trace_rays(rays, crt.RAY_CLOSEST)
    hit, hit_ref = face >= 0, t_ref < 1e30
    assert (hit != hit_ref).mean() < 2e-4
    both = hit & hit_ref
    rel = np.abs(t[both] - t_ref[both]) / t_ref[both]
    print("%s: hit/miss disagreements %d of %d; max rel t diff %.3g; exact %.4f" % (name, int((hit != hit_ref).sum()), len(rays), rel.max(), (rel == 0).mean()))
    assert np.quantile(rel, 0.999) < 1e-5
