"""GPU tests on the procedural workloads C4 / C5 (tools/synthetic.py): oracle parity at sizes the CPU finishes in
seconds, size-independent properties at the full 10M-triangle size."""
import math

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def gpu(crt):
    if crt.device_count() < 1:
        pytest.fail("no CUDA device")
    return crt


def test_c4_small_parity(gpu, orc):
    from tools import synthetic as sy
    verts, mat, obj, mats = sy.c4_scene(97)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2)
    b = orc.Scene().add_arrays(verts, mat.astype(np.int32), obj.astype(np.int32), mats)
    onodes, oorder, olast, obounds = b.build_new_bvh(2)
    nodes, order, last, bounds = a.export_bvh()
    assert nodes.tobytes() == onodes.tobytes() and np.array_equal(order, oorder) and np.array_equal(last, olast)
    c = sy.C4_CAMERA
    M = gpu.inverse_view_matrix(c["eye"], c["lookat"], c["up"])
    W, H, spp = 192, 108, 3
    R = gpu.Render(a, W, H, spp, c["P_RR"], c["light_sample_n"])
    R.run_view(c["eye"], M, math.radians(c["fov_y"]))
    oacc, ost = b.render(c["eye"], M, math.radians(c["fov_y"]), W, H, 0, spp, c["P_RR"], c["light_sample_n"])
    assert np.array_equal(R.get_accum_i64(), oacc)
    assert R.stats()["shadow_rays"] == ost["shadow_rays"]


def test_c5_random_rays_match_numpy_statement_and_oracle(gpu, orc):
    import torch
    from tools import synthetic as sy
    verts, mat, obj, mats = sy.c4_scene(65)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2)
    b = orc.Scene().add_arrays(verts, mat.astype(np.int32), obj.astype(np.int32), mats)
    b.build_new_bvh(2)
    _, _, _, bounds = a.export_bvh()
    n = 20000
    for any_hit in (False, True):
        d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        a.random_rays_device(d_rays.data_ptr(), n, start=12345, key=0xC5, any_hit=any_hit)
        torch.cuda.synchronize()
        rays = d_rays.cpu().numpy()
        want = sy.random_rays(bounds[:3], bounds[3:], n, key=0xC5, any_hit=any_hit, start=12345)
        assert np.array_equal(rays.view(np.uint32), want.view(np.uint32))
        d_t = torch.empty(n, dtype=torch.float32, device="cuda")
        d_f = torch.empty(n, dtype=torch.int32, device="cuda")
        mode = gpu.RAY_ANY if any_hit else gpu.RAY_CLOSEST
        a.trace_rays_device(d_rays.data_ptr(), n, mode, d_t.data_ptr(), d_f.data_ptr())
        if any_hit:
            assert b.check_any_hits(rays, d_t.cpu().numpy(), d_f.cpu().numpy())
            continue
        ot, of = b.trace(rays, which=0, mode=mode)
        assert np.array_equal(d_f.cpu().numpy(), of) and np.array_equal(d_t.cpu().numpy().view(np.uint32), ot.view(np.uint32))
    assert abs(np.linalg.norm(rays[:, 4:7], axis=1) - 1).max() < 1e-6


def test_c4_full_size_properties(gpu):
    """10M triangles: the BVH is a partition of the triangles into sentinel-terminated leaves of <= thresh
    triangles, boxes nest, tracing is deterministic, any-hit agrees with closest-hit, shards add up exactly."""
    import torch
    from tools import synthetic as sy
    verts, mat, obj, mats = sy.c4_scene(sy.C4_FULL_N)
    n = len(verts)
    assert n == 2 * (sy.C4_FULL_N - 1) ** 2 + 12
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    ms = a.set_BVH(2)
    nodes, order, last, bounds = a.export_bvh()
    assert np.array_equal(np.sort(order), np.arange(n, dtype=np.int32))
    assert nodes["n0"][0] + nodes["n1"][0] == n
    leaf0, leaf1 = nodes["c0"] < 0, nodes["c1"] < 0
    assert (nodes["n0"][leaf0] <= 2).all() and (nodes["n1"][leaf1] <= 2).all()
    assert int(nodes["n0"][leaf0].sum() + nodes["n1"][leaf1].sum()) == n
    assert int(last.sum()) == int(leaf0.sum() + leaf1.sum())
    # child counts are consistent with the children's own counts; child boxes nest in the parent's
    for c, cnt, lo, hi in (("c0", "n0", ("c0lox", "c0loy", "c0loz"), ("c0hix", "c0hiy", "c0hiz")),
                           ("c1", "n1", ("c1lox", "c1loy", "c1loz"), ("c1hix", "c1hiy", "c1hiz"))):
        inner = nodes[c] >= 0
        ch = nodes[c][inner]
        assert np.array_equal(nodes[cnt][inner], nodes["n0"][ch] + nodes["n1"][ch])
        for k in range(3):
            child_lo = np.minimum(nodes[("c0lox", "c0loy", "c0loz")[k]][ch], nodes[("c1lox", "c1loy", "c1loz")[k]][ch])
            child_hi = np.maximum(nodes[("c0hix", "c0hiy", "c0hiz")[k]][ch], nodes[("c1hix", "c1hiy", "c1hiz")[k]][ch])
            assert np.array_equal(nodes[lo[k]][inner], child_lo) and np.array_equal(nodes[hi[k]][inner], child_hi)
    assert np.array_equal(bounds, np.concatenate([verts.reshape(-1, 3).min(0), verts.reshape(-1, 3).max(0)]))
    # rays
    nr = 2_000_000
    d_rays = torch.empty((nr, 8), dtype=torch.float32, device="cuda")
    d_t, d_f = torch.empty(nr, dtype=torch.float32, device="cuda"), torch.empty(nr, dtype=torch.int32, device="cuda")
    d_t2, d_f2 = torch.empty_like(d_t), torch.empty_like(d_f)
    a.random_rays_device(d_rays.data_ptr(), nr, any_hit=True)
    a.trace_rays_device(d_rays.data_ptr(), nr, gpu.RAY_CLOSEST, d_t.data_ptr(), d_f.data_ptr())
    a.trace_rays_device(d_rays.data_ptr(), nr, gpu.RAY_CLOSEST, d_t2.data_ptr(), d_f2.data_ptr())
    assert torch.equal(d_t, d_t2) and torch.equal(d_f, d_f2)
    a.trace_rays_device(d_rays.data_ptr(), nr, gpu.RAY_ANY, d_t2.data_ptr(), d_f2.data_ptr())
    tmax = d_rays[:, 3]
    blocked_by_closest = (d_f >= 0) & ((tmax - d_t) > 1e-5)
    assert torch.equal(blocked_by_closest, d_f2 >= 0)         # exists a blocker <=> the closest hit is a blocker
    # brute-force spot check of 64 rays against all 10M triangles on the GPU (torch, float64 tolerance on t)
    idx = torch.arange(0, nr, nr // 64, device="cuda")[:64]
    V = torch.from_numpy(verts).cuda().double().view(-1, 3, 3)
    for i in idx.tolist():
        o, d = d_rays[i, 0:3].double(), d_rays[i, 4:7].double()
        e1, e2 = V[:, 1] - V[:, 0], V[:, 2] - V[:, 0]
        s = o - V[:, 0]
        s1, s2 = torch.linalg.cross(d.expand_as(e2), e2), torch.linalg.cross(s, e1)
        rcp = 1.0 / (s1 * e1).sum(1)
        bb, gg, tt = (s1 * s).sum(1) * rcp, (s2 * d).sum(1) * rcp, (s2 * e2).sum(1) * rcp
        ok = (bb > 1e-9) & (gg > 1e-9) & (bb + gg < 1 - 1e-9) & (tt > 1e-4)
        tmin = tt[ok].min().item() if ok.any() else float("inf")
        got = d_t[i].item() if d_f[i].item() >= 0 else float("inf")
        assert (math.isinf(tmin) and math.isinf(got)) or abs(got - tmin) <= 1e-4 * max(1.0, tmin), (i, got, tmin)
    # render: two half-shards add up to the whole
    c = sy.C4_CAMERA
    M = gpu.inverse_view_matrix(c["eye"], c["lookat"], c["up"])
    R = gpu.Render(a, 480, 270, 4, c["P_RR"], c["light_sample_n"])
    R.run_view(c["eye"], M, math.radians(c["fov_y"]))
    full = R.get_accum_i64()
    R.set_sample_range(0, 1); R.run_view(c["eye"], M, math.radians(c["fov_y"])); part = R.get_accum_i64()
    R.set_sample_range(1, 4); R.run_view(c["eye"], M, math.radians(c["fov_y"])); part = part + R.get_accum_i64()
    assert np.array_equal(full, part) and full.any()
    print("C4 full: %d triangles, %d nodes, GPU build %.1f ms" % (n, len(nodes), ms))


def test_c4_full_size_rays_equal_the_oracle(gpu, orc):
    """The full 10M-triangle scene, builder ploc8 (the bench default): 200k of the C5 rays, closest-hit (t bits, face id)
    and any-hit decisions, against the oracle's traversal of its own CPU build of the same tree; and a crop of the C4
    frame (work range of the sample-major index space) against the oracle's buffer rows."""
    import torch
    from tools import synthetic as sy
    verts, mat, obj, mats = sy.c4_scene(sy.C4_FULL_N)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2, builder=3)
    b = orc.Scene().add_arrays(verts, mat.astype(np.int32), obj.astype(np.int32), mats)
    b.build_new_bvh(2, 2)
    b.build_wide8(2, 3)
    n = 200000
    for any_hit in (False, True):
        d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
        a.random_rays_device(d_rays.data_ptr(), n, start=0, key=0xC5, any_hit=any_hit)
        d_t, d_f = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
        mode = gpu.RAY_ANY if any_hit else gpu.RAY_CLOSEST
        a.trace_rays_device(d_rays.data_ptr(), n, mode, d_t.data_ptr(), d_f.data_ptr())
        rays, t, f = d_rays.cpu().numpy(), d_t.cpu().numpy(), d_f.cpu().numpy()
        if any_hit:
            assert b.check_any_hits(rays, t, f) and 0.3 < (f >= 0).mean() < 0.9
        else:
            ot, of = b.trace(rays, which=4, mode=0)
            assert np.array_equal(f, of) and np.array_equal(t.view(np.uint32), ot.view(np.uint32)) and (f >= 0).mean() > 0.5
            ht, hf, _ = a.trace_rays(rays, gpu.RAY_CLOSEST)            # the host-buffer entry point (chunked pipeline) gives the same
            assert np.array_equal(hf, of) and np.array_equal(ht.view(np.uint32), ot.view(np.uint32))
    c = sy.C4_CAMERA
    M = gpu.inverse_view_matrix(c["eye"], c["lookat"], c["up"])
    W, H = 480, 270
    R = gpu.Render(a, W, H, 2, c["P_RR"], c["light_sample_n"])
    R.run_view(c["eye"], M, math.radians(c["fov_y"]))
    oacc, ost = b.render(c["eye"], M, math.radians(c["fov_y"]), W, H, 0, 2, c["P_RR"], c["light_sample_n"], wide=True)
    assert np.array_equal(R.get_accum_i64(), oacc) and oacc.any()


def test_host_buffer_batches_are_chunked_and_exact(gpu, orc, monkeypatch):
    """crt_trace_rays with host buffers runs in chunks on two streams (CRT_BATCH_CHUNK rays each): many small chunks, a
    ragged last one, pageable and page-locked caller buffers all give the device-buffer result."""
    import torch
    from tools import synthetic as sy
    monkeypatch.setenv("CRT_BATCH_CHUNK", "4096")
    verts, mat, obj, mats = sy.c4_scene(65)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2, builder=3)
    n = 4096 * 5 + 123
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    a.random_rays_device(d_rays.data_ptr(), n, start=7, key=0xC5, any_hit=True)
    d_t, d_f = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    rays = d_rays.cpu().numpy()
    for mode in (gpu.RAY_CLOSEST, gpu.RAY_ANY):
        a.trace_rays_device(d_rays.data_ptr(), n, mode, d_t.data_ptr(), d_f.data_ptr())
        want_t, want_f = d_t.cpu().numpy(), d_f.cpu().numpy()
        t, f, ms = a.trace_rays(rays, mode)                               # pageable in, pageable out
        assert ms > 0 and np.array_equal(f >= 0, want_f >= 0)
        if mode == gpu.RAY_CLOSEST:
            assert np.array_equal(f, want_f) and np.array_equal(t.view(np.uint32), want_t.view(np.uint32))
        pin_rays = torch.from_numpy(rays).pin_memory()
        pin_t, pin_f = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
        t2, f2, _ = a.trace_rays(pin_rays.numpy(), mode, out=(pin_t.numpy(), pin_f.numpy()))     # page-locked: DMA in place
        assert np.array_equal(f2 >= 0, want_f >= 0)
        if mode == gpu.RAY_CLOSEST:
            assert np.array_equal(f2, want_f) and np.array_equal(t2.view(np.uint32), want_t.view(np.uint32))
    t, f, _ = a.trace_rays(rays[:0], gpu.RAY_CLOSEST)
    assert len(t) == 0


def test_default_chunks_page_locked_then_pageable_on_one_scene(gpu, monkeypatch):
    """Without CRT_BATCH_CHUNK page-locked caller buffers go in half-size chunks and pageable ones in whole chunks through the same
    staging buffers of the scene: mixed kinds of buffers, one scene, more rays than one chunk, the same hits every way."""
    import torch
    from tools import synthetic as sy
    monkeypatch.delenv("CRT_BATCH_CHUNK", raising=False)
    verts, mat, obj, mats = sy.c4_scene(33)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2, builder=3)
    n = (4 << 20) + (2 << 20) + 4321                                      # one whole chunk and a ragged rest; three half chunks
    d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
    a.random_rays_device(d_rays.data_ptr(), n, start=11, key=0xC5, any_hit=False)
    d_t, d_f = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
    a.trace_rays_device(d_rays.data_ptr(), n, gpu.RAY_CLOSEST, d_t.data_ptr(), d_f.data_ptr())
    want_t, want_f = d_t.cpu().numpy(), d_f.cpu().numpy()
    pin_rays = d_rays.cpu().pin_memory()
    pageable = np.array(pin_rays.numpy())
    pin_t, pin_f = torch.empty(n, dtype=torch.float32).pin_memory(), torch.empty(n, dtype=torch.int32).pin_memory()
    # page-locked rays with pageable results first (staging for the results is sized here), then everything pageable, then all page-locked
    for rays, out in ((pin_rays.numpy(), None), (pageable, None), (pin_rays.numpy(), (pin_t.numpy(), pin_f.numpy())), (pageable, (pin_t.numpy(), pin_f.numpy()))):
        t, f, _ = a.trace_rays(rays, gpu.RAY_CLOSEST, out=out) if out is not None else a.trace_rays(rays, gpu.RAY_CLOSEST)
        assert np.array_equal(f, want_f) and np.array_equal(t.view(np.uint32), want_t.view(np.uint32))


def test_sorted_batches_give_the_same_hits(gpu, monkeypatch):
    """CRT_RAY_SORTED: the batch is traced in (origin cell, direction octant) order and the hits come back in the caller's order -
    the same hits as without the flag, through the device-buffer call, the chunked host-buffer call (both pipeline slots, a ragged
    last chunk) and for a batch smaller than one sort tile."""
    import torch
    from tools import synthetic as sy
    monkeypatch.setenv("CRT_BATCH_CHUNK", "8192")
    verts, mat, obj, mats = sy.c4_scene(65)
    a = gpu.Scene().add_triangles(verts, mat, obj, mats)
    a.set_BVH(2, builder=3)
    for n in (8192 * 4 + 777, 300):
        for any_hit, mode in ((False, gpu.RAY_CLOSEST), (True, gpu.RAY_ANY)):
            d_rays = torch.empty((n, 8), dtype=torch.float32, device="cuda")
            a.random_rays_device(d_rays.data_ptr(), n, start=3, key=0xC5, any_hit=any_hit)
            t0, f0 = torch.empty(n, dtype=torch.float32, device="cuda"), torch.empty(n, dtype=torch.int32, device="cuda")
            t1, f1 = torch.full((n,), -7.0, dtype=torch.float32, device="cuda"), torch.full((n,), -7, dtype=torch.int32, device="cuda")
            a.trace_rays_device(d_rays.data_ptr(), n, mode, t0.data_ptr(), f0.data_ptr())
            ms = a.trace_rays_device(d_rays.data_ptr(), n, mode | gpu.RAY_SORTED, t1.data_ptr(), f1.data_ptr())
            assert ms > 0
            t0h, f0h, t1h, f1h = t0.cpu().numpy(), f0.cpu().numpy(), t1.cpu().numpy(), f1.cpu().numpy()
            assert (f0h >= 0).any() and (f0h < 0).any()
            if mode == gpu.RAY_CLOSEST:
                assert np.array_equal(f1h, f0h) and np.array_equal(t1h.view(np.uint32), t0h.view(np.uint32))
            else:
                assert np.array_equal(f1h >= 0, f0h >= 0)
            ht, hf, _ = a.trace_rays(d_rays.cpu().numpy(), mode | gpu.RAY_SORTED)          # chunks of 8192 on two streams
            if mode == gpu.RAY_CLOSEST:
                assert np.array_equal(hf, f0h) and np.array_equal(ht.view(np.uint32), t0h.view(np.uint32))
            else:
                assert np.array_equal(hf >= 0, f0h >= 0)
    with pytest.raises(gpu.CrtError):
        a.trace_rays(np.zeros((4, 8), np.float32), 0x200)
