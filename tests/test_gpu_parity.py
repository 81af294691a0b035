"""GPU parity tests (run on the B200 box with -m gpu). Everything goes through the C-ABI (libcrt.so) and is
compared with the CPU oracle: BVH bytes, hit ids / t bits, the fixed-point accumulation buffer."""
import math
import os

import numpy as np
import pytest

from conftest import BOX_CAMERA, box_scene, random_rays, soup

pytestmark = pytest.mark.gpu

FLT_MAX = np.finfo(np.float32).max


@pytest.fixture(scope="module")
def gpu(crt):
    if crt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback and these tests need the B200")
    return crt


def _pair(gpu, orc, files, name, thresh=None):
    cfg = gpu.load_config(files[name]["cfg_path"])
    th = thresh if thresh is not None else cfg.bvh_thresh_n
    a = gpu.Scene().add_obj(files[name]["obj"], files[name]["dir"])
    a.set_BVH(th)
    b = orc.Scene().add_obj(files[name]["obj"], files[name]["dir"])
    b.build_new_bvh(th)
    return cfg, a, b


def _same_bvh(a, b_built):
    nodes, order, last, bounds = a.export_bvh()
    onodes, oorder, olast, obounds = b_built
    assert len(nodes) == len(onodes)
    assert nodes.tobytes() == onodes.tobytes(), "node bytes differ"
    assert np.array_equal(order, oorder) and np.array_equal(last, olast) and np.array_equal(bounds, obounds)


@pytest.mark.parametrize("name,thresh", [("veach-mis", 2), ("veach-mis", 1), ("veach-mis", 6), ("cornell-box", 2), ("cornell-box", 4)])
def test_gpu_bvh_build_is_bit_exact(gpu, orc, scene_files, name, thresh):
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(thresh)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    _same_bvh(a, b.build_new_bvh(thresh))


@pytest.mark.parametrize("n,thresh", [(1, 1), (1, 4), (2, 1), (2, 2), (3, 1), (5, 8), (1024, 2), (1025, 2), (4097, 3), (50000, 2), (300000, 4)])
def test_gpu_bvh_build_random_soups(gpu, orc, n, thresh):
    rng = np.random.default_rng(n * 31 + thresh)
    verts = soup(rng, n, extent=20.0, size=0.7)
    if n > 100:
        verts[10:40] = verts[10]                       # duplicate triangles -> duplicate Morton keys
        verts[50:60, [1, 4, 7]] = 3.0                  # flat, axis-aligned boxes
    mats = [[.5, .5, .5, 0, 0, 0, 1]]
    a = gpu.Scene().add_triangles(verts, np.zeros(n, np.uint32), np.zeros(n, np.uint32), mats)
    a.set_BVH(thresh)
    b = orc.Scene().add_arrays(verts, np.zeros(n, np.int32), np.zeros(n, np.int32), mats)
    _same_bvh(a, b.build_new_bvh(thresh))
    rays = random_rays(rng, [-22] * 3, [22] * 3, 20000)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))


def test_gpu_bvh_flat_and_degenerate_extent(gpu, orc):
    """All triangles in one plane (zero extent on an axis) and all-identical triangles."""
    rng = np.random.default_rng(11)
    verts = soup(rng, 3000, extent=5.0, size=0.5)
    verts[:, [2, 5, 8]] = 1.25
    same = np.repeat(soup(rng, 1), 257, axis=0)
    for v in (verts, same):
        n = len(v)
        a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), [[.5, .5, .5, 0, 0, 0, 1]])
        a.set_BVH(2)
        b = orc.Scene().add_arrays(v, np.zeros(n, np.int32), np.zeros(n, np.int32), [[.5, .5, .5, 0, 0, 0, 1]])
        _same_bvh(a, b.build_new_bvh(2))


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_primary_hit_ids_bit_exact_vs_reference_rule(gpu, orc, scene_files, name):
    """North-star check 1: primary-ray hit triangle ids from the GPU (new BVH, new traversal) equal the
    reference's host BVH + traversal rule (oracle transcription, canonical ties), at the shipped 800x600."""
    cfg, a, b = _pair(gpu, orc, scene_files, name)
    b.build_ref_bvh(cfg.bvh_thresh_n)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    rays = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), cfg.width, cfg.height)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    t_ref, f_ref = b.trace(rays, which=1)
    assert np.array_equal(f, f_ref)
    assert np.array_equal(t.view(np.uint32), t_ref.view(np.uint32))
    gold = np.load(os.path.join(os.path.dirname(__file__), "golden", name + "_hits_200x150.npz"))
    rays_s = orc.primary_rays(cfg.eye_pos, M, float(cfg.fovy_rad), 200, 150)
    t_s, f_s, _ = a.trace_rays(rays_s, gpu.RAY_CLOSEST)
    assert np.array_equal(f_s, gold["face"]) and np.array_equal(t_s.view(np.uint32), gold["t_bits"])


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_incoherent_rays_closest_and_any(gpu, orc, scene_files, name):
    cfg, a, b = _pair(gpu, orc, scene_files, name)
    _, _, _, bounds = a.export_bvh()
    rng = np.random.default_rng(2)
    for mode in (gpu.RAY_CLOSEST, gpu.RAY_ANY):
        rays = random_rays(rng, bounds[:3], bounds[3:], 300000, tmax_any=(mode == gpu.RAY_ANY))
        t, f, _ = a.trace_rays(rays, mode)
        if mode == gpu.RAY_ANY:
            assert b.check_any_hits(rays, t, f)               # "a blocker or none": which one is scheduling-dependent
            continue
        ot, of = b.trace(rays, which=0, mode=mode)
        assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    # brute force on a sample: the traversal never loses the closest triangle
    rays = random_rays(rng, bounds[:3], bounds[3:], 300 if name == "cornell-box" else 5000)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    bt, bf = b.trace(rays, which=3, mode=0)
    assert np.array_equal(f, bf) and np.array_equal(t.view(np.uint32), bt.view(np.uint32))


def test_ray_edge_cases(gpu, orc):
    verts, mat, obj, mats = box_scene(np.random.default_rng(0), 100)
    a = gpu.Scene().add_triangles(verts, mat.astype(np.uint32), obj.astype(np.uint32), mats)
    a.set_BVH(2)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    b.build_new_bvh(2)
    rays = np.array([
        [5, 5, 5, FLT_MAX, 1, 0, 0, 0], [5, 5, 5, FLT_MAX, 0, -1, 0, 0], [5, 5, 5, FLT_MAX, 0, 0, 1, 0],      # axis-aligned (inf inv_dir)
        [5, 5, 5, FLT_MAX, -1, 0, 0, 0], [5, 0, 5, FLT_MAX, 1, 0, 0, 0],                                       # in the floor plane
        [0, 0, 0, FLT_MAX, 0.57735026, 0.57735026, 0.57735026, 0],                                               # through a corner
        [5, 5, -30, FLT_MAX, 0, 0, -1, 0],                                                                       # away from everything
        [5, 5, 5, 1e-3, 0, 1, 0, 0], [5, 5, 5, 0.0, 0, 1, 0, 0],                                                 # tiny / zero tmax
    ], np.float32)
    t, f, _ = a.trace_rays(rays, 0)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    t, f, _ = a.trace_rays(rays, 1)
    assert b.check_any_hits(rays, t, f)
    t, f, ms = a.trace_rays(np.zeros((0, 8), np.float32), 0)                                                     # empty batch
    assert len(t) == 0 and len(f) == 0


def _render_pair(gpu, orc, a, b, cam_eye, M, fovy, W, H, spp, p_rr, lsn, seed=0):
    R = gpu.Render(a, W, H, spp, p_rr, lsn)
    R.set_seed(seed)
    R.run_view(cam_eye, M, fovy)
    acc = R.get_accum_i64()
    oacc, ost = b.render(cam_eye, M, float(fovy), W, H, 0, spp, p_rr, lsn, seed=seed)
    return R, acc, oacc, ost


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_matched_seed_radiance_shipped_configs(gpu, orc, scene_files, name):
    """North-star check 2 at BASELINE configs C1 / C2 (full 800x600, shipped spp): radiance under matched
    seeds. Stated tolerance 1e-4 relative per pixel; the fixed-point buffers are in fact identical."""
    cfg, a, b = _pair(gpu, orc, scene_files, name)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R, acc, oacc, ost = _render_pair(gpu, orc, a, b, cfg.eye_pos, M, cfg.fovy_rad, cfg.width, cfg.height, cfg.spp, cfg.P_RR, cfg.light_sample_n)
    st = R.stats()
    assert (st["extend_rays"], st["shadow_rays"], st["probe_rays"]) == (ost["extend_rays"], ost["shadow_rays"], ost["probe_rays"])
    lin = R.get_accum().reshape(-1)
    olin, orgb = orc.resolve(oacc, cfg.width * cfg.height, cfg.spp)
    rel = np.abs(lin - olin) / np.maximum(np.abs(olin), 1e-6)
    assert rel.max() <= 1e-4                                  # the north star's tolerance
    assert np.array_equal(acc, oacc)                          # and what actually holds: exact equality
    rgb = R.get_frame_buffer().reshape(-1)
    assert np.abs(rgb.astype(int) - orgb.astype(int)).max() <= 1     # powf on device vs host: at most one code value
    assert (rgb != orgb).mean() < 1e-3


def test_matched_seed_radiance_synthetic_specular_and_options(gpu, orc):
    verts, mat, obj, mats = box_scene(np.random.default_rng(1), 400)
    a = gpu.Scene().add_triangles(verts, mat.astype(np.uint32), obj.astype(np.uint32), mats)
    a.set_BVH(2)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    b.build_new_bvh(2)
    M = gpu.inverse_view_matrix(BOX_CAMERA["eye"], BOX_CAMERA["lookat"], BOX_CAMERA["up"])
    for (W, H, spp, p_rr, lsn, seed) in [(96, 64, 8, 0.6, 2, 0), (33, 17, 3, 0.9, 1, 5), (64, 64, 2, 1.0, 3, 9), (50, 40, 4, 0.0, 1, 2)]:
        R, acc, oacc, ost = _render_pair(gpu, orc, a, b, BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"], W, H, spp, p_rr, lsn, seed)
        assert np.array_equal(acc, oacc), (W, H, spp, p_rr, lsn, seed)
        assert R.stats()["probe_rays"] == ost["probe_rays"]
        if p_rr > 0.5:
            assert ost["probe_rays"] > 0                      # the SPECULAR plate is exercised
    # P_RR = 1: paths stop at 64 vertices (BOUNCE_STACK_SIZE)
    assert R.stats()["iterations"] <= 66


def test_scheduling_knobs_do_not_change_the_image(gpu, orc, scene_files, monkeypatch):
    """k_shadow on a second stream beside the next iteration's extend (CRT_OVERLAP), the pool size and the hand-over to
    k_tail only reorder work: the integer accumulation buffer and the ray counts stay the same."""
    cfg, a, b = _pair(gpu, orc, scene_files, "cornell-box")
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 256, 192, 5
    want = None
    for env in ({}, {"CRT_OVERLAP": "0"}, {"CRT_OVERLAP": "1", "CRT_POOL": "8192"}, {"CRT_OVERLAP": "1", "CRT_TAIL": "0"},
                {"CRT_OVERLAP": "0", "CRT_POOL": "4096", "CRT_TAIL": "1000000"}):
        for k in ("CRT_OVERLAP", "CRT_POOL", "CRT_TAIL"):
            monkeypatch.delenv(k, raising=False)
        for k, v in env.items():
            monkeypatch.setenv(k, v)
        R = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
        R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
        got, st = R.get_accum_i64().copy(), R.stats()
        if want is None:
            want, want_st = got, st
            oacc, ost = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 0, spp, cfg.P_RR, cfg.light_sample_n)
            assert np.array_equal(want, oacc) and st["shadow_rays"] == ost["shadow_rays"]
        assert np.array_equal(got, want), env
        assert (st["extend_rays"], st["shadow_rays"]) == (want_st["extend_rays"], want_st["shadow_rays"]), env


def test_render_is_deterministic_and_shards_add_up(gpu, orc, scene_files):
    """Multi-GPU invariance on one GPU: any split of the work index space sums to the full buffer exactly,
    and repeated runs are identical (atomics are integer, so ordering cannot matter)."""
    cfg, a, b = _pair(gpu, orc, scene_files, "veach-mis")
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 320, 240, 6
    R = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    full = R.get_accum_i64()
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    assert np.array_equal(full, R.get_accum_i64())
    from cudaraytracing_b200.distributed import shard_work
    for world in (2, 4, 8):
        total = np.zeros_like(full)
        for r in range(world):
            w0, w1 = shard_work(W * H, spp, r, world)
            R.set_work_range(w0, w1)
            R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
            total += R.get_accum_i64()
        assert np.array_equal(total, full), world
    R.set_sample_range(2, 5)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, _ = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 2, 5, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(R.get_accum_i64(), oacc)
    R.clear_range()
    R.set_spp(1)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, _ = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 0, 1, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(R.get_accum_i64(), oacc)


def test_pool_smaller_than_frame(gpu, orc, scene_files, monkeypatch):
    """Path regeneration: a pool far smaller than the number of samples gives the same image."""
    monkeypatch.setenv("CRT_POOL", "8192")
    cfg, a, b = _pair(gpu, orc, scene_files, "veach-mis")
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R, acc, oacc, _ = _render_pair(gpu, orc, a, b, cfg.eye_pos, M, cfg.fovy_rad, 200, 150, 3, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(acc, oacc)
    assert R.stats()["iterations"] > 12


@pytest.mark.parametrize("tail", ["0", "1", "3000", "1000000"])
def test_tail_switch_point_does_not_change_the_image(gpu, orc, scene_files, monkeypatch, tail):
    """k_tail (remaining paths finished in place, one lane per path) may take over at any point of the
    frame: never, for the last path only, mid-way, or straight after the first wavefront."""
    monkeypatch.setenv("CRT_TAIL", tail)
    cfg, a, b = _pair(gpu, orc, scene_files, "veach-mis")
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R, acc, oacc, ost = _render_pair(gpu, orc, a, b, cfg.eye_pos, M, cfg.fovy_rad, 160, 120, 3, 0.8, 2)
    st = R.stats()
    assert np.array_equal(acc, oacc)
    assert (st["extend_rays"], st["shadow_rays"], st["probe_rays"]) == (ost["extend_rays"], ost["shadow_rays"], ost["probe_rays"])
    if tail == "1000000":
        assert st["iterations"] == 2                          # one wavefront for the camera rays, then the tail
    if tail == "0":
        assert st["iterations"] > 8


def _mis_pair(gpu, orc, a, b, eye, M, fovy, W, H, spp, p_rr, lsn, seed=0):
    R = gpu.Render(a, W, H, spp, p_rr, lsn)
    R.set_seed(seed)
    R.set_estimator(gpu.ESTIMATOR_MIS)
    R.run_view(eye, M, fovy)
    oacc, ost = b.render(eye, M, float(fovy), W, H, 0, spp, p_rr, lsn, seed=seed, estimator=1)
    return R, R.get_accum_i64(), oacc, ost


@pytest.mark.parametrize("name,W,H", [("veach-mis", 800, 600), ("cornell-box", 400, 300)])
def test_mis_estimator_matched_seed(gpu, orc, scene_files, name, W, H):
    """North star (c): NEE with MIS, Phong lobes from Kd/Ks/Ns. C2 (veach-mis) at its full shipped size; the
    fixed-point buffers are identical to the CPU statement (stated tolerance: 1e-4 relative per pixel)."""
    cfg, a, b = _pair(gpu, orc, scene_files, name)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    R, acc, oacc, ost = _mis_pair(gpu, orc, a, b, cfg.eye_pos, M, cfg.fovy_rad, W, H, cfg.spp, cfg.P_RR, cfg.light_sample_n)
    st = R.stats()
    assert (st["extend_rays"], st["shadow_rays"], st["probe_rays"]) == (ost["extend_rays"], ost["shadow_rays"], 0)
    lin = R.get_accum().reshape(-1)
    olin, _ = orc.resolve(oacc, W * H, cfg.spp)
    assert (np.abs(lin - olin) / np.maximum(np.abs(olin), 1e-6)).max() <= 1e-4
    assert np.array_equal(acc, oacc)
    assert acc.any()


@pytest.mark.parametrize("tail", ["0", "2000", "1000000"])
def test_mis_estimator_options_and_tail(gpu, orc, monkeypatch, tail):
    monkeypatch.setenv("CRT_TAIL", tail)
    verts, mat, obj, mats7 = box_scene(np.random.default_rng(3), 300)
    mats = np.zeros((len(mats7), 10), np.float32)                       # (kd, ks, ke, ns)
    mats[:, 0:3] = mats7[:, 0:3]; mats[:, 6:9] = mats7[:, 3:6]; mats[:, 9] = mats7[:, 6]
    mats[4, 3:6] = [0.5, 0.4, 0.3]; mats[4, 9] = 60.0                    # glossy plate
    mats[5, 3:6] = [0.2, 0.2, 0.2]; mats[5, 9] = 8.0                     # glossy soup
    a = gpu.Scene().add_triangles(verts, mat.astype(np.uint32), obj.astype(np.uint32), mats)
    a.set_BVH(2)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    b.build_new_bvh(2)
    M = gpu.inverse_view_matrix(BOX_CAMERA["eye"], BOX_CAMERA["lookat"], BOX_CAMERA["up"])
    for (W, H, spp, p_rr, lsn, seed) in [(96, 64, 6, 0.7, 2, 0), (33, 17, 3, 0.95, 1, 5), (64, 48, 2, 1.0, 3, 9), (40, 30, 4, 0.0, 1, 2)]:
        R, acc, oacc, ost = _mis_pair(gpu, orc, a, b, BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"], W, H, spp, p_rr, lsn, seed)
        assert np.array_equal(acc, oacc), (W, H, spp, p_rr, lsn, seed)
        assert R.stats()["shadow_rays"] == ost["shadow_rays"] and ost["shadow_rays"] > 0
    # shards of the work index space add up for mis too
    R = gpu.Render(a, 64, 48, 4, 0.7, 2)
    R.set_estimator(gpu.ESTIMATOR_MIS)
    R.run_view(BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"])
    full = R.get_accum_i64()
    from cudaraytracing_b200.distributed import shard_work
    total = np.zeros_like(full)
    for r in range(3):
        w0, w1 = shard_work(64 * 48, 4, r, 3)
        R.set_work_range(w0, w1)
        R.run_view(BOX_CAMERA["eye"], M, BOX_CAMERA["fovy"])
        total += R.get_accum_i64()
    assert np.array_equal(total, full)


@pytest.mark.parametrize("builder", [0, 1, 2, 3])
@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_rays_lying_in_box_planes(gpu, orc, scene_files, name, builder):
    """Rays from surface points with one direction component exactly zero (they lie in box planes): the committed
    regression rays and fresh ones, both node layouts, against the oracle (which equals brute force on them,
    tests/test_oracle.py::test_rays_lying_in_box_planes_equal_brute_force)."""
    from conftest import GOLDEN
    from test_oracle import _surface_rays_with_zero_components
    cfg = gpu.load_config(scene_files[name]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(cfg.bvh_thresh_n, builder=builder)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    b.build_new_bvh(cfg.bvh_thresh_n)
    z = np.load(os.path.join(GOLDEN, "axis_planar_rays.npz"))
    batches = [(z[k], 1 if "_any_" in k else 0) for k in z.files if k.startswith(name)]
    rng = np.random.default_rng(23)
    fresh = _surface_rays_with_zero_components(orc, b, cfg, rng, n_max=40000)
    batches.append((fresh, 0))
    anyr = fresh.copy()
    anyr[:, 3] = rng.uniform(0, 900.0 if name == "cornell-box" else 30.0, len(anyr)).astype(np.float32)
    batches.append((anyr, 1))
    # rays exactly along a coordinate axis (two zero components): must be culled on both parallel axes - the kernel time
    # bound below fails when such rays walk the whole tree (35 ms per ray on cornell-box, profiles/r01_s15.md)
    lo, hi = b.build_new_bvh(cfg.bvh_thresh_n)[3][:3], b.build_new_bvh(cfg.bvh_thresh_n)[3][3:]
    n = 100000
    ax = np.zeros((n, 8), np.float32)
    ax[:, 0:3] = rng.uniform(lo, hi, (n, 3))
    ax[np.arange(n), 4 + rng.integers(0, 3, n)] = rng.choice([-1.0, 1.0], n)
    ax[:, 3] = FLT_MAX
    t, f, ms = a.trace_rays(ax, 0)
    ot, of = b.trace(ax, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    assert ms < 20.0, "axis-parallel rays took %.1f ms" % ms
    if name == "veach-mis":      # the shadow ray through a box corner (tests/test_oracle.py::test_shadow_ray_through_a_box_corner)
        corner = np.array([[-5, 1.35669553, 2.52801561, 8.02512932, 0.570246458, 0.669184923, -0.476456344, 0]], np.float32)
        t, f, _ = a.trace_rays(corner, 1)
        assert f[0] == 2022
    for rays, mode in batches:
        t, f, _ = a.trace_rays(rays, mode)
        if mode == 0:
            ot, of = b.trace(rays, which=0, mode=0)
            assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
        else:
            assert b.check_any_hits(rays, t, f)


def test_progressive_chunks_and_checkpoint_resume(gpu, scene_files, tmp_path):
    """a frame rendered in chunks (accumulate on), interrupted, saved, loaded into a fresh handle and
    finished equals the frame rendered in one go, bit for bit; mismatching settings / camera / corrupt files are refused."""
    name = "veach-mis"
    cfg = gpu.load_config(scene_files[name]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(cfg.bvh_thresh_n)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 200, 150, 6
    full = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    full.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    want = full.get_accum_i64()
    png_full = str(tmp_path / "full.png")
    full.save_frame_buffer(png_full)
    npix = W * H
    ck = str(tmp_path / "render.ckpt")
    # first process: 2 chunks of ragged size, then "crash"
    r1 = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    r1.set_accumulate(True)
    r1.clear_accum()
    done = 0
    for end in (npix + 777, 3 * npix + 5):
        r1.set_work_range(done, end)
        r1.run_view(cfg.eye_pos, M, cfg.fovy_rad)
        done = end
        r1.save_checkpoint(ck, done)
    with pytest.raises(gpu.CrtError) as e:        # accumulate on + another camera
        r1.run_view(np.asarray(cfg.eye_pos) + 1.0, M, cfg.fovy_rad)
    assert e.value.code == -5
    del r1
    # second process: resume
    r2 = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    wd, eye, Mc, fov = r2.load_checkpoint(ck)
    assert wd == done and np.array_equal(eye, np.asarray(cfg.eye_pos, np.float32)) and np.array_equal(Mc.reshape(9), np.asarray(M, np.float32).reshape(9))
    assert np.float32(fov) == np.float32(cfg.fovy_rad)
    r2.set_work_range(wd, npix * spp)
    r2.run_view(eye, Mc, fov)
    assert np.array_equal(r2.get_accum_i64(), want)
    png2 = str(tmp_path / "resumed.png")
    r2.save_frame_buffer(png2)
    assert open(png2, "rb").read() == open(png_full, "rb").read()
    # refused: other settings, truncated file, flipped payload byte, not a checkpoint
    r3 = gpu.Render(a, W, H, spp + 1, cfg.P_RR, cfg.light_sample_n)
    with pytest.raises(gpu.CrtError) as e:
        r3.load_checkpoint(ck)
    assert e.value.code == -5
    data = open(ck, "rb").read()
    bad = str(tmp_path / "bad.ckpt")
    for blob in (data[:-9], data[:200] + bytes([data[200] ^ 1]) + data[201:], b"not a checkpoint" * 10, data + b"x"):
        open(bad, "wb").write(blob)
        with pytest.raises(gpu.CrtError) as e:
            gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n).load_checkpoint(bad)
        assert e.value.code == -2
    # refused: a render of another scene (the header carries the scene's identity), a work_done beyond the frame
    other = gpu.Scene().add_obj(scene_files["cornell-box"]["obj"], scene_files["cornell-box"]["dir"])
    other.set_BVH(2)
    with pytest.raises(gpu.CrtError) as e:
        gpu.Render(other, W, H, spp, cfg.P_RR, cfg.light_sample_n).load_checkpoint(ck)
    assert e.value.code == -5 and "different scene" in str(e.value)
    with pytest.raises(gpu.CrtError):
        r2.save_checkpoint(str(tmp_path / "over.ckpt"), W * H * spp + 1)
    with pytest.raises(gpu.CrtError) as e:
        gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n).save_checkpoint(ck, 0)      # nothing rendered yet
    assert e.value.code == -5


def test_cli_checkpoint_resume(gpu, scene_files, tmp_path):
    import json
    import subprocess
    from conftest import ROOT
    exe = os.path.join(ROOT, "cudaraytracing_b200", "crt")
    base = [exe, "--config", scene_files["veach-mis"]["cfg_path"], "--width", "160", "--height", "120", "--spp", "7"]
    one = str(tmp_path / "one.png")
    r = subprocess.run(base + ["--out", one], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    ck, two = str(tmp_path / "c.ckpt"), str(tmp_path / "two.png")
    r = subprocess.run(base + ["--out", two, "--checkpoint", ck, "--chunk-spp", "2", "--stop-after", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["complete"] is False and info["work_done"] == 4 * 160 * 120 and not os.path.exists(two)
    r = subprocess.run(base + ["--out", two, "--checkpoint", ck, "--chunk-spp", "2"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["resumed_from"] == 4 * 160 * 120
    assert open(one, "rb").read() == open(two, "rb").read()
    r = subprocess.run(base + ["--out", two, "--checkpoint", ck, "--seed", "5"], capture_output=True, text=True)
    assert r.returncode == 1 and "different render settings" in r.stderr


def test_save_png_and_cli(gpu, scene_files, tmp_path):
    import subprocess
    import struct
    import zlib
    from conftest import ROOT
    exe = os.path.join(ROOT, "cudaraytracing_b200", "crt")
    out = str(tmp_path / "v.png")
    r = subprocess.run([exe, "--config", scene_files["veach-mis"]["cfg_path"], "--out", out, "--width", "160", "--height", "120"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    import json
    info = json.loads(r.stdout.strip().splitlines()[-1])
    assert info["triangles"] == 3092 and info["lights"] == 4
    data = open(out, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n" and struct.unpack(">II", data[16:24]) == (160, 120)
    # same image through the API
    cfg = gpu.load_config(scene_files["veach-mis"]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files["veach-mis"]["obj"], scene_files["veach-mis"]["dir"])
    a.set_BVH(cfg.bvh_thresh_n)
    R = gpu.Render(a, 160, 120, cfg.spp, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up), cfg.fovy_rad)
    p2 = str(tmp_path / "api.png")
    R.save_frame_buffer(p2)
    assert open(p2, "rb").read() == data


def test_error_paths_on_gpu(gpu):
    s = gpu.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [0], [0], [[.5, .5, .5, 0, 0, 0, 1]])
    with pytest.raises(gpu.CrtError) as e:
        gpu.Render(s, 8, 8)
    assert e.value.code == -5
    with pytest.raises(gpu.CrtError):
        s.set_BVH(2, device=99)
    s.set_BVH(2)
    with pytest.raises(gpu.CrtError) as e:
        s.add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [0], [0], [[.5, .5, .5, 0, 0, 0, 1]])
    assert e.value.code == -5
    R = gpu.Render(s, 8, 8)
    with pytest.raises(gpu.CrtError):
        R.set_P_RR(1.5)
    with pytest.raises(gpu.CrtError):
        R.set_estimator(7)
    # a scene with no lights and an empty scene still render (to black)
    R.run_view([0.2, 0.2, -1], np.eye(3, dtype=np.float32).reshape(9), 0.5)
    assert not R.get_accum_i64().any()
    e0 = gpu.Scene()
    e0.set_BVH(2)
    R0 = gpu.Render(e0, 4, 4)
    R0.run_view([0, 0, 0], np.eye(3, dtype=np.float32).reshape(9), 0.5)
    assert not R0.get_accum_i64().any()


def test_c3_size_frame_equals_the_oracle(gpu, orc, scene_files, monkeypatch):
    """The benchmarked configuration at its own size: cornell-box 3840x2160, builder ploc8, 2^25 paths in flight (the pool
    a C3 frame runs with), camera paths started tile by tile (whole samples, frame larger than a tile) - two samples per
    pixel of it, every value of the 199 MB fixed-point buffer equal to the oracle's."""
    monkeypatch.setenv("CRT_POOL", str(1 << 25))
    cfg = gpu.load_config(scene_files["cornell-box"]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files["cornell-box"]["obj"], scene_files["cornell-box"]["dir"])
    a.set_BVH(cfg.bvh_thresh_n, builder=3)
    b = orc.Scene().add_obj(scene_files["cornell-box"]["obj"], scene_files["cornell-box"]["dir"])
    b.build_new_bvh(cfg.bvh_thresh_n, 2)
    b.build_wide8(cfg.bvh_thresh_n, 3)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 3840, 2160, 2
    R = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    acc, st = R.get_accum_i64(), R.stats()
    oacc, ost = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 0, spp, cfg.P_RR, cfg.light_sample_n, wide=True)
    assert np.array_equal(acc, oacc), "%d values differ" % int((acc != oacc).sum())
    assert (st["extend_rays"], st["shadow_rays"]) == (ost["extend_rays"], ost["shadow_rays"]) and st["samples"] == W * H * spp
    # a band of rows of one sample (a work range that is not whole samples: sample-major order) adds up with the rest
    R.set_work_range(1000 * W, 1064 * W); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); band = R.get_accum_i64()
    assert band.reshape(H, W, 3)[:1000].any() == False and band.reshape(H, W, 3)[1064:].any() == False and band.any()
    R.set_work_range(0, 1000 * W); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); band += R.get_accum_i64()
    R.set_work_range(1064 * W, W * H * spp); R.run_view(cfg.eye_pos, M, cfg.fovy_rad); band += R.get_accum_i64()
    assert np.array_equal(band, oacc)


@pytest.mark.parametrize("tile_px", ["0", "1000", "4096", "47999", "48000"])
def test_tile_order_does_not_change_the_buffer(gpu, orc, scene_files, monkeypatch, tile_px):
    """k_generate may start the work items of whole-sample ranges tile by tile (CRT_TILE_PX pixels per tile, last tile
    ragged): any order gives the same integer buffer."""
    monkeypatch.setenv("CRT_TILE_PX", tile_px)
    monkeypatch.setenv("CRT_POOL", "16384")
    cfg, a, b = _pair(gpu, orc, scene_files, "veach-mis")
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 240, 200, 3
    R = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, ost = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 0, spp, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(R.get_accum_i64(), oacc) and R.stats()["extend_rays"] == ost["extend_rays"]
    R.set_sample_range(1, 3)
    R.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, _ = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 1, 3, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(R.get_accum_i64(), oacc)
