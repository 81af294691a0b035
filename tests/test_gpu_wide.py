"""8-wide compressed BVH (builder LBVH8) on the GPU: the node bytes equal the CPU statement's
(oracle build_wide8_bvh), and hits / images equal the pair-node BVH's — conservative boxes make the
result independent of the tree, so every oracle answer computed on the pair-node BVH applies."""
import numpy as np
import pytest

from conftest import BOX_CAMERA, box_scene, random_rays, soup

pytestmark = pytest.mark.gpu

MAT = [[.5, .5, .5, 0, 0, 0, 1]]


@pytest.fixture(scope="module")
def gpu(crt):
    if crt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback and these tests need the B200")
    return crt


def _same_wide(gpu, a, built):
    assert a.bvh_kind() == gpu.BUILDER_LBVH8
    nodes, order, last, bounds = a.export_bvh()
    onodes, oorder, olast, obounds = built
    assert len(nodes) == len(onodes)
    assert nodes.tobytes() == onodes.tobytes(), "wide node bytes differ"
    assert np.array_equal(order, oorder) and np.array_equal(last, olast) and np.array_equal(bounds, obounds)


@pytest.mark.parametrize("name,thresh", [("veach-mis", 2), ("veach-mis", 1), ("veach-mis", 6), ("cornell-box", 2), ("cornell-box", 4),
                                         ("cornell-box", 40)])
def test_wide_build_is_bit_exact(gpu, orc, scene_files, name, thresh):
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(thresh, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    _same_wide(gpu, a, b.build_wide8(thresh))


@pytest.mark.parametrize("n,thresh", [(1, 1), (1, 4), (2, 1), (2, 2), (3, 1), (5, 8), (9, 1), (1024, 2), (1025, 2), (4097, 3), (50000, 2),
                                      (300000, 4)])
def test_wide_build_and_trace_random_soups(gpu, orc, n, thresh):
    rng = np.random.default_rng(n * 17 + thresh)
    verts = soup(rng, n, extent=20.0, size=0.7)
    if n > 100:
        verts[10:40] = verts[10]                       # duplicate triangles -> duplicate Morton keys
        verts[50:60, [1, 4, 7]] = 3.0                  # flat, axis-aligned boxes
    a = gpu.Scene().add_triangles(verts, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
    a.set_BVH(thresh, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_arrays(verts, np.zeros(n, np.int32), np.zeros(n, np.int32), MAT)
    _same_wide(gpu, a, b.build_wide8(thresh))
    b.build_new_bvh(thresh)
    rays = random_rays(rng, [-22] * 3, [22] * 3, 20000)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    rays = random_rays(rng, [-22] * 3, [22] * 3, 20000, tmax_any=True)
    t, f, _ = a.trace_rays(rays, gpu.RAY_ANY)
    assert b.check_any_hits(rays, t, f)


def test_wide_flat_and_degenerate_extent(gpu, orc):
    rng = np.random.default_rng(11)
    verts = soup(rng, 3000, extent=5.0, size=0.5)
    verts[:, [2, 5, 8]] = 1.25                         # everything in one plane: zero extent on z
    same = np.repeat(soup(rng, 1), 257, axis=0)        # identical triangles: a chain of duplicate keys
    for v in (verts, same):
        n = len(v)
        a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
        a.set_BVH(2, builder=gpu.BUILDER_LBVH8)
        b = orc.Scene().add_arrays(v, np.zeros(n, np.int32), np.zeros(n, np.int32), MAT)
        _same_wide(gpu, a, b.build_wide8(2))
        b.build_new_bvh(2)
        rays = random_rays(rng, [-6] * 3, [6] * 3, 5000)
        t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
        ot, of = b.trace(rays, which=0, mode=0)
        assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_wide_incoherent_rays_and_primary_hits(gpu, orc, scene_files, name):
    cfg = gpu.load_config(scene_files[name]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(cfg.bvh_thresh_n, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    b.build_new_bvh(cfg.bvh_thresh_n)
    b.build_ref_bvh(cfg.bvh_thresh_n)
    _, _, _, bounds = a.export_bvh()
    rng = np.random.default_rng(4)
    rays = random_rays(rng, bounds[:3], bounds[3:], 300000)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    rays = random_rays(rng, bounds[:3], bounds[3:], 300000, tmax_any=True)
    t, f, _ = a.trace_rays(rays, gpu.RAY_ANY)
    assert b.check_any_hits(rays, t, f)
    # primary rays at the shipped size against the reference's host BVH + traversal rule
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    t_ref, f_ref = b.trace(rays, which=1)
    assert np.array_equal(f, f_ref) and np.array_equal(t.view(np.uint32), t_ref.view(np.uint32))


def test_wide_ray_edge_cases(gpu, orc):
    FLT_MAX = np.finfo(np.float32).max
    verts, mat, obj, mats = box_scene(np.random.default_rng(0), 100)
    a = gpu.Scene().add_triangles(verts, mat.astype(np.uint32), obj.astype(np.uint32), mats)
    a.set_BVH(2, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    b.build_new_bvh(2)
    rays = np.array([
        [5, 5, 5, FLT_MAX, 1, 0, 0, 0], [5, 5, 5, FLT_MAX, 0, -1, 0, 0], [5, 5, 5, FLT_MAX, 0, 0, 1, 0],
        [5, 5, 5, FLT_MAX, -1, 0, 0, 0], [5, 0, 5, FLT_MAX, 1, 0, 0, 0],
        [0, 0, 0, FLT_MAX, 0.57735026, 0.57735026, 0.57735026, 0],
        [5, 5, -30, FLT_MAX, 0, 0, -1, 0],
        [5, 5, 5, 1e-3, 0, 1, 0, 0], [5, 5, 5, 0.0, 0, 1, 0, 0],
    ], np.float32)
    t, f, _ = a.trace_rays(rays, 0)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    t, f, _ = a.trace_rays(rays, 1)
    assert b.check_any_hits(rays, t, f)
    e0 = gpu.Scene()
    e0.set_BVH(2, builder=gpu.BUILDER_LBVH8)                    # empty scene
    t, f, _ = e0.trace_rays(rays, 0)
    assert np.all(f == -1)


@pytest.mark.parametrize("name,est", [("cornell-box", 0), ("veach-mis", 0), ("veach-mis", 1)])
def test_wide_render_equals_the_oracle_image(gpu, orc, scene_files, name, est):
    """Shipped configs C1 / C2 on the wide BVH: the fixed-point accumulation buffer equals the oracle's."""
    cfg = gpu.load_config(scene_files[name]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(cfg.bvh_thresh_n, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    b.build_new_bvh(cfg.bvh_thresh_n)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R = gpu.Render(a, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
    R.set_estimator(est)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, ost = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height, 0, cfg.spp, cfg.P_RR, cfg.light_sample_n, estimator=est)
    st = R.stats()
    assert (st["extend_rays"], st["shadow_rays"], st["probe_rays"]) == (ost["extend_rays"], ost["shadow_rays"], ost["probe_rays"])
    assert np.array_equal(R.get_accum_i64(), oacc)


@pytest.mark.parametrize("tail", ["0", "3000", "1000000"])
def test_wide_render_specular_scene_and_tail(gpu, orc, monkeypatch, tail):
    monkeypatch.setenv("CRT_TAIL", tail)
    verts, mat, obj, mats = box_scene(np.random.default_rng(1), 400)
    a = gpu.Scene().add_triangles(verts, mat.astype(np.uint32), obj.astype(np.uint32), mats)
    a.set_BVH(2, builder=gpu.BUILDER_LBVH8)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    b.build_new_bvh(2)
    M = gpu.inverse_view_matrix(BOX_CAMERA["eye"], BOX_CAMERA["lookat"], BOX_CAMERA["up"])
    for (W, H, spp, p_rr, lsn, seed) in [(96, 64, 8, 0.6, 2, 0), (33, 17, 3, 0.9, 1, 5)]:
        R = gpu.Render(a, W, H, spp, p_rr, lsn)
        R.set_seed(seed)
        R.run_view(BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"])
        oacc, ost = b.render(BOX_CAMERA["eye"], M, float(BOX_CAMERA["fovy"]), W, H, 0, spp, p_rr, lsn, seed=seed)
        assert np.array_equal(R.get_accum_i64(), oacc), (W, H, spp, p_rr, lsn, seed)
        assert R.stats()["probe_rays"] == ost["probe_rays"] and ost["probe_rays"] > 0


def test_export_refuses_the_wrong_layout(gpu):
    import ctypes as C
    s = gpu.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [0], [0], MAT)
    s.set_BVH(2, builder=gpu.BUILDER_LBVH8)
    assert s.L.crt_scene_export_bvh(s.h, None, None, None, None) == -5
    s2 = gpu.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [0], [0], MAT)
    s2.set_BVH(2)
    assert s2.L.crt_scene_export_bvh8(s2.h, None, None, None, None) == -5
    with pytest.raises(gpu.CrtError):
        s2.set_BVH(2, builder=7)
