"""Pins the oracle's restatement against the REAL reference code (oracle/_ref/libref.so = the reference's
own headers compiled in place). Needs /root/reference, so it runs in the build container only."""
import ctypes as C
import os

import numpy as np
import pytest

from conftest import ROOT, REFERENCE

pytestmark = pytest.mark.reference


@pytest.fixture(scope="module")
def ref():
    so = os.path.join(ROOT, "oracle", "_ref", "libref.so")
    if not os.path.exists(so):
        import subprocess
        subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_harness", "build.sh")])
    R = C.CDLL(so)
    R.ref_host_load.restype = C.c_void_p
    R.ref_host_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint]
    for f in ("ref_n_tris", "ref_n_nodes", "ref_root", "ref_n_lights"):
        getattr(R, f).argtypes = [C.c_void_p]
    R.ref_get_tris.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    R.ref_get_nodes.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_get_light.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    R.ref_inverse_view.argtypes = [C.c_void_p] * 4
    return R


@pytest.mark.parametrize("name,thresh", [("veach-mis", 2), ("veach-mis", 5), ("cornell-box", 2)])
def test_oracle_equals_real_reference_host_path(ref, orc, crt, name, thresh):
    d = "%s/scenes/%s/" % (REFERENCE, name)
    h = ref.ref_host_load((d + name + ".obj").encode(), d.encode(), 800, 600, thresh)
    n, nn = ref.ref_n_tris(h), ref.ref_n_nodes(h)
    t0 = np.zeros((n, 23), np.float32); ref.ref_get_tris(h, 0, t0.ctypes.data_as(C.c_void_p))
    t1 = np.zeros((n, 23), np.float32); ref.ref_get_tris(h, 1, t1.ctypes.data_as(C.c_void_p))
    nodes = np.zeros(nn, orc.REF_NODE); ref.ref_get_nodes(h, nodes.ctypes.data_as(C.c_void_p))
    for S in (orc.Scene().add_obj(d + name + ".obj", d), crt.Scene().add_obj(d + name + ".obj", d)):
        tr, mats = S.tris(), S.mats()
        bits = lambda a: np.ascontiguousarray(a).view(np.uint32)
        assert np.array_equal(bits(tr["verts"]), bits(t0[:, 0:9]))
        assert np.array_equal(bits(tr["normal"]), bits(t0[:, 9:12]))
        assert np.array_equal(bits(tr["area"]), bits(t0[:, 12]))
        assert np.array_equal(bits(tr["area_of_obj"]), bits(t0[:, 13]))
        m = mats[tr["mat"]]
        assert np.array_equal(m[:, 0:3], t0[:, 14:17]) and np.array_equal(m[:, 3:6], t0[:, 17:20])
        assert np.array_equal(m[:, 6], t0[:, 20]) and np.array_equal(m[:, 7], t0[:, 21]) and np.array_equal(m[:, 8], t0[:, 22])
        for li, (faces, area) in enumerate(S.lights()):
            nt, ar = C.c_int(), C.c_float()
            ref.ref_get_light(h, li, C.byref(nt), C.byref(ar))
            assert nt.value == len(faces) and np.float32(ar.value) == np.float32(area)
    S = orc.Scene().add_obj(d + name + ".obj", d)
    mynodes, order, root = S.build_ref_bvh(thresh)
    assert root == ref.ref_root(h) and mynodes.tobytes() == nodes.tobytes()          # BVH.h:37-84, byte for byte
    assert np.array_equal(S.tris()["verts"][order].view(np.uint32), np.ascontiguousarray(t1[:, 0:9]).view(np.uint32))


def test_camera_matrix_equals_reference(ref, orc, crt):
    for eye, look, up in ([[278, 273, -800], [278, 273, -799], [0, 1, 0]], [[28.2792, 5.2, 1.23612e-06], [0, 2.8, 0], [0, 1, 0]]):
        e, l, u = (np.asarray(x, np.float32) for x in (eye, look, up))
        out = np.zeros(9, np.float32)
        ref.ref_inverse_view(*(a.ctypes.data_as(C.c_void_p) for a in (e, l, u, out)))
        assert np.array_equal(out, orc.inverse_view_matrix(eye, look, up))
        assert np.array_equal(out, crt.inverse_view_matrix(eye, look, up))


def test_fixture_equals_reference_scene_files(crt, scene_files):
    for name in ("cornell-box", "veach-mis"):
        d = "%s/scenes/%s/" % (REFERENCE, name)
        a = crt.Scene().add_obj(d + name + ".obj", d)
        b = crt.Scene().add_obj(scene_files[name]["obj"], scene_files[name]["dir"])
        ta, tb = a.tris(), b.tris()
        for k in ta:
            assert np.array_equal(ta[k].view(np.uint32), tb[k].view(np.uint32)), k
        assert np.array_equal(a.mats(), b.mats())
        ca, cb = crt.load_config(d + "config.json"), crt.load_config(scene_files[name]["cfg_path"])
        for k in ("fov_y", "width", "height", "bvh_thresh_n", "P_RR", "spp", "light_sample_n"):
            assert getattr(ca, k) == getattr(cb, k)
        assert np.array_equal(ca.eye_pos, cb.eye_pos) and np.array_equal(ca.lookat, cb.lookat)
