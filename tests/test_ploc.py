"""PLOC topology (builders PLOC = 2, PLOC8 = 3): bottom-up agglomeration over the Morton order.

CPU part: the oracle's statement (oracle/orc_bvh.cpp build_ploc_tree) yields a valid tree, the same hits as the
Karras tree and brute force (conservative boxes make results tree-independent), fewer node visits, and a bounded
number of rounds on degenerate input. GPU part (-m gpu): the CUDA builder reproduces node bytes / triangle order /
leaf terminators of that statement exactly, for both node layouts, and renders the same fixed-point image."""
import numpy as np
import pytest

from conftest import BOX_CAMERA, box_scene, random_rays, soup

MAT = [[.5, .5, .5, 0, 0, 0, 1]]
PLOC, PLOC8 = 2, 3


def _scene(orc, verts):
    n = len(verts)
    return orc.Scene().add_arrays(verts, np.zeros(n, np.int32), np.zeros(n, np.int32), MAT)


def _check_pair_tree(nodes, order, last, bounds, verts, thresh):
    n = len(verts)
    assert sorted(order.tolist()) == list(range(n))
    tri = verts.reshape(n, 3, 3)
    tlo, thi = tri.min(axis=1), tri.max(axis=1)
    seen = np.zeros(n, np.int32)
    referenced = np.zeros(len(nodes), np.int32)
    referenced[0] = 1
    sub_lo = np.zeros((len(nodes), 3), np.float32)
    sub_hi = np.zeros((len(nodes), 3), np.float32)
    cnt_of = np.zeros(len(nodes), np.int64)
    for i in range(len(nodes) - 1, -1, -1):                 # children have larger indices than their parent
        nd = nodes[i]
        los, his = [], []
        for w in (0, 1):
            c, cnt = int(nd["c%d" % w]), int(nd["n%d" % w])
            blo = np.array([nd["c%dlox" % w], nd["c%dloy" % w], nd["c%dloz" % w]])
            bhi = np.array([nd["c%dhix" % w], nd["c%dhiy" % w], nd["c%dhiz" % w]])
            if c < 0:
                first = ~c
                assert 1 <= cnt <= thresh
                assert last[first + cnt - 1] == 1 and not last[first:first + cnt - 1].any()
                seen[first:first + cnt] += 1
                faces = order[first:first + cnt]
                clo, chi = tlo[faces].min(axis=0), thi[faces].max(axis=0)
            else:
                assert c > i and cnt == cnt_of[c] and cnt > thresh
                referenced[c] += 1
                clo, chi = sub_lo[c], sub_hi[c]
            assert np.array_equal(blo, clo) and np.array_equal(bhi, chi), "child box is not the exact bound of its content"
            los.append(clo); his.append(chi)
        sub_lo[i] = np.minimum(*los); sub_hi[i] = np.maximum(*his)
        cnt_of[i] = int(nd["n0"]) + int(nd["n1"])
    assert (seen == 1).all() and (referenced == 1).all() and cnt_of[0] == n
    assert np.array_equal(sub_lo[0], bounds[:3]) and np.array_equal(sub_hi[0], bounds[3:])


def test_ploc_structure(orc):
    rng = np.random.default_rng(3)
    n = 2000
    verts = soup(rng, n)
    verts[100:140] = verts[100]                              # identical triangles: every union area ties
    verts[300:360, [1, 4, 7]] = 2.0                          # flat boxes
    S = _scene(orc, verts)
    for thresh in (1, 2, 3, 8, 64):
        nodes, order, last, bounds = S.build_new_bvh(thresh, PLOC)
        _check_pair_tree(nodes, order, last, bounds, verts, thresh)


def test_ploc_hits_equal_lbvh_and_brute_force(orc, scene_files):
    rng = np.random.default_rng(9)
    verts = soup(rng, 3000)
    verts[:200, [2, 5, 8]] = 0.5
    S = _scene(orc, verts)
    for thresh in (1, 2, 5):
        for any_mode in (0, 1):
            rays = random_rays(rng, [-12] * 3, [12] * 3, 4000, tmax_any=bool(any_mode))
            rays[:50, 4:7] = [1, 0, 0]
            rays[50:100, 4:7] = [0, 0, -1]
            t3, f3 = S.trace(rays, which=3, mode=any_mode)
            S.build_new_bvh(thresh, PLOC)
            t2, f2 = S.trace(rays, which=0, mode=any_mode)
            S.build_wide8(thresh, PLOC8)
            t4, f4 = S.trace(rays, which=4, mode=any_mode)
            if any_mode == 0:
                for t, f in ((t2, f2), (t4, f4)):
                    assert np.array_equal(f, f3) and np.array_equal(t.view(np.uint32), t3.view(np.uint32))
            else:
                assert np.array_equal(f2 >= 0, f3 >= 0) and np.array_equal(f4 >= 0, f3 >= 0)


@pytest.mark.parametrize("name,gain", [("cornell-box", 0.75), ("veach-mis", 0.85)])
def test_ploc_needs_fewer_node_visits_on_the_shipped_scenes(orc, crt, scene_files, name, gain):
    """The reason the builder exists: path-tracing rays (primary + bounce + shadow) of the shipped configs visit
    clearly fewer nodes than with the Karras tree, the image being bit-identical."""
    cfg = crt.load_config(scene_files[name]["cfg_path"])
    S = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    M = orc.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    res = {}
    for b in (0, PLOC):
        S.build_new_bvh(cfg.bvh_thresh_n, b)
        res[b] = S.render(cfg.eye_pos, M, float(cfg.fovy_rad), 120, 90, 0, 2, cfg.P_RR, cfg.light_sample_n)
    assert np.array_equal(res[0][0], res[PLOC][0])
    s0, s2 = res[0][1], res[PLOC][1]
    assert s2["closest_inner"] < gain * s0["closest_inner"]
    assert s2["any_inner"] < s0["any_inner"]
    # and through the wide layout
    S.build_wide8(cfg.bvh_thresh_n, PLOC8)
    acc, s3 = S.render(cfg.eye_pos, M, float(cfg.fovy_rad), 120, 90, 0, 2, cfg.P_RR, cfg.light_sample_n, wide=True)
    assert np.array_equal(acc, res[0][0])


def test_ploc_single_leaf_scene_keeps_the_morton_order(orc):
    """n <= thresh_n: the scene is one leaf, no topology is built, and every builder emits the Morton order
    (first GPU run of the PLOC builder, session 16: the oracle permuted the slots of a 5-triangle leaf)."""
    rng = np.random.default_rng(5 * 13 + 8)
    verts = soup(rng, 5, extent=20.0, size=0.7)
    S = _scene(orc, verts)
    ref = S.build_new_bvh(8, 0)
    for b in (PLOC, PLOC8):
        got = S.build_wide8(8, b) if b & 1 else S.build_new_bvh(8, b)
        assert np.array_equal(got[1], ref[1]) and np.array_equal(got[2], ref[2])
    assert ref[0].tobytes() == S.build_new_bvh(8, PLOC)[0].tobytes()


def test_ploc_degenerate_inputs(orc):
    one = np.array([[0, 0, 0, 1, 0, 0, 0, 1, 0]], np.float32)
    for n in (1, 2, 3, 9):
        verts = np.repeat(one, n, axis=0) + np.arange(n, dtype=np.float32)[:, None]
        S = _scene(orc, verts)
        for thresh in (1, 2, 5):
            rays = np.array([[0.2 + (n - 1), 0.2 + (n - 1), -1 + (n - 1), 3e38, 0, 0, 1, 0]], np.float32)
            tb, fb = S.trace(rays, which=3)
            nodes, order, last, bounds = S.build_new_bvh(thresh, PLOC)
            assert len(nodes) >= 1 and last[-1] == 1
            if n > thresh:
                _check_pair_tree(nodes, order, last, bounds, verts, thresh)
            t, f = S.trace(rays, which=0)
            assert f[0] == fb[0] == n - 1 and t[0] == tb[0]
            S.build_wide8(thresh, PLOC8)
            t, f = S.trace(rays, which=4)
            assert f[0] == n - 1 and t[0] == tb[0]
    # identical triangles: all areas tie; the buddy rule pairs them up, so the tree stays shallow
    same = np.repeat(one, 1000, axis=0)
    S = _scene(orc, same)
    nodes, order, last, bounds = S.build_new_bvh(2, PLOC)
    _check_pair_tree(nodes, order, last, bounds, same, 2)
    t, f, st = S.trace(np.array([[0.2, 0.2, -1, 3e38, 0, 0, 1, 0]], np.float32), which=0, want_stats=True)
    assert f[0] == 0 and t[0] == 1.0 and st["max_stack"] <= 12
    S = orc.Scene()
    assert len(S.build_new_bvh(2, PLOC)[0]) == 0               # empty scene
    assert S.trace(np.array([[0, 0, 0, 3e38, 0, 0, 1, 0]], np.float32), which=0)[1][0] == -1


# ---------------------------------------------------------------------------------------------------- GPU
@pytest.fixture(scope="module")
def gpu(crt):
    if crt.device_count() < 1:
        pytest.fail("no CUDA device: the product has no CPU fallback and these tests need the B200")
    return crt


def _same_build(gpu, a, b, thresh, builder):
    a.set_BVH(thresh, builder=builder)
    assert a.bvh_kind() == builder
    nodes, order, last, bounds = a.export_bvh()
    onodes, oorder, olast, obounds = b.build_wide8(thresh, builder) if builder & 1 else b.build_new_bvh(thresh, builder)
    assert len(nodes) == len(onodes)
    assert np.array_equal(order, oorder), "triangle order differs"
    assert nodes.tobytes() == onodes.tobytes(), "node bytes differ"
    assert np.array_equal(last, olast) and np.array_equal(bounds, obounds)


@pytest.mark.gpu
@pytest.mark.parametrize("builder", [PLOC, PLOC8])
@pytest.mark.parametrize("name,thresh", [("veach-mis", 2), ("veach-mis", 1), ("veach-mis", 6), ("cornell-box", 2), ("cornell-box", 4)])
def test_gpu_ploc_build_is_bit_exact(gpu, orc, scene_files, name, thresh, builder):
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    _same_build(gpu, a, b, thresh, builder)


@pytest.mark.gpu
@pytest.mark.parametrize("builder", [PLOC, PLOC8])
@pytest.mark.parametrize("n,thresh", [(1, 1), (1, 4), (2, 1), (2, 2), (3, 1), (5, 8), (9, 1), (255, 2), (256, 2), (257, 2), (1024, 2), (1025, 2),
                                      (4097, 3), (50000, 2), (300000, 4)])
def test_gpu_ploc_build_and_trace_random_soups(gpu, orc, n, thresh, builder):
    rng = np.random.default_rng(n * 13 + thresh)
    verts = soup(rng, n, extent=20.0, size=0.7)
    if n > 100:
        verts[10:40] = verts[10]                           # identical triangles: ties everywhere
        verts[50:60, [1, 4, 7]] = 3.0                      # flat, axis-aligned boxes
    a = gpu.Scene().add_triangles(verts, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
    b = _scene(orc, verts)
    _same_build(gpu, a, b, thresh, builder)
    b.build_new_bvh(thresh)
    rays = random_rays(rng, [-22] * 3, [22] * 3, 20000)
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    ot, of = b.trace(rays, which=0, mode=0)
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32))
    rays = random_rays(rng, [-22] * 3, [22] * 3, 20000, tmax_any=True)
    t, f, _ = a.trace_rays(rays, gpu.RAY_ANY)
    assert b.check_any_hits(rays, t, f)


@pytest.mark.gpu
@pytest.mark.parametrize("builder", [PLOC, PLOC8])
def test_gpu_ploc_flat_and_identical(gpu, orc, builder):
    rng = np.random.default_rng(11)
    flat = soup(rng, 3000, extent=5.0, size=0.5)
    flat[:, [2, 5, 8]] = 1.25
    same = np.repeat(soup(rng, 1), 2000, axis=0)
    for v in (flat, same):
        n = len(v)
        a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
        _same_build(gpu, a, _scene(orc, v), 2, builder)


@pytest.mark.gpu
@pytest.mark.parametrize("builder", [PLOC, PLOC8])
@pytest.mark.parametrize("name,est", [("cornell-box", 0), ("veach-mis", 0), ("veach-mis", 1)])
def test_gpu_ploc_render_equals_the_oracle_image(gpu, orc, scene_files, name, est, builder):
    cfg = gpu.load_config(scene_files[name]["cfg_path"])
    a = gpu.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    a.set_BVH(cfg.bvh_thresh_n, builder=builder)
    b = orc.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
    b.build_new_bvh(cfg.bvh_thresh_n)
    M = gpu.inverse_view_matrix(cfg.eye_pos, cfg.lookat, cfg.up)
    W, H, spp = 400, 300, 2
    r = gpu.Render(a, W, H, spp, cfg.P_RR, cfg.light_sample_n)
    r.set_estimator(est)
    r.run_view(cfg.eye_pos, M, cfg.fovy_rad)
    oacc, ost = b.render(cfg.eye_pos, M, float(cfg.fovy_rad), W, H, 0, spp, cfg.P_RR, cfg.light_sample_n, estimator=est)
    acc = r.get_accum_i64()
    assert np.array_equal(acc, oacc), "%d values differ" % int((acc != oacc).sum())
    st = r.stats()
    assert st["extend_rays"] == ost["extend_rays"] and st["shadow_rays"] == ost["shadow_rays"]


@pytest.mark.gpu
def test_gpu_ploc_synthetic_heightfield(gpu, orc):
    """The C4 generator at a reduced grid: build parity on a regular tessellation (many exactly equal areas)."""
    from tools import synthetic as sy
    v, m, o, mats = sy.c4_scene(129)
    a = gpu.Scene().add_triangles(v, m, o, mats)
    b = orc.Scene().add_arrays(v, m.astype(np.int32), o.astype(np.int32), mats)
    for builder in (PLOC, PLOC8):
        _same_build(gpu, a if builder == PLOC else gpu.Scene().add_triangles(v, m, o, mats), b, 2, builder)


@pytest.mark.gpu
def test_gpu_ploc_chain_is_refused_not_overflowed(gpu, orc):
    """Boxes that grow geometrically merge one pair per round: a chain as deep as the triangle count. The pair-node
    traversal stack is kStackSize = 96 entries without a bounds check, so builder PLOC refuses such a tree with a
    message (it used to overflow the stack); PLOC8 collapses the chain into shallow wide levels and traces it."""
    n = 200
    x = (1.5 ** np.arange(n)).astype(np.float32)
    v = np.zeros((n, 9), np.float32)
    v[:, 0] = x; v[:, 3] = x * 1.01; v[:, 6] = x; v[:, 7] = x * 0.01 + 1e-3
    a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
    with pytest.raises(gpu.CrtError) as e:
        a.set_BVH(1, builder=PLOC)
    assert "more rounds than the traversal stack" in str(e.value)
    a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
    b = _scene(orc, v)
    _same_build(gpu, a, b, 1, PLOC8)
    rays = np.zeros((n, 8), np.float32)
    rays[:, 0] = x * 1.002; rays[:, 1] = (x * 0.01 + 1e-3) * 0.25; rays[:, 2] = -1.0; rays[:, 3] = 3e38; rays[:, 6] = 1.0
    t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    ot, of = b.trace(rays, which=3)                            # brute force
    assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and (f >= 0).sum() > n // 2
    a = gpu.Scene().add_triangles(v, np.zeros(n, np.uint32), np.zeros(n, np.uint32), MAT)
    a.set_BVH(1, builder=0)                                   # the Karras tree of the same scene is shallow enough
    t2, f2, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)
    assert np.array_equal(f2, of)
