"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/liborc.so (the CPU restatement of the reference's algorithm and of the
new builder / traversal rule / estimators). Imported only by tests/, __graft_entry__.smoke() and
bench.py's cpu_baseline / --impl reference legs; the product package never imports it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

f32p = np.ctypeslib.ndpointer(np.float32, flags="C_CONTIGUOUS")
i32p = np.ctypeslib.ndpointer(np.int32, flags="C_CONTIGUOUS")
u8p = np.ctypeslib.ndpointer(np.uint8, flags="C_CONTIGUOUS")
i64p = np.ctypeslib.ndpointer(np.int64, flags="C_CONTIGUOUS")
u64p = np.ctypeslib.ndpointer(np.uint64, flags="C_CONTIGUOUS")
u32p = np.ctypeslib.ndpointer(np.uint32, flags="C_CONTIGUOUS")


def build(force=False):
    so = os.path.join(_HERE, "liborc.so")
    srcs = [os.path.join(_HERE, f) for f in os.listdir(_HERE) if f.endswith((".cpp", ".h"))]
    if force or not os.path.exists(so) or any(os.path.getmtime(s) > os.path.getmtime(so) for s in srcs):
        subprocess.check_call(["make", "-C", _HERE, "liborc.so"], stdout=subprocess.DEVNULL)
    return so


def lib():
    global _LIB
    if _LIB is not None:
        return _LIB
    L = C.CDLL(build())
    vp = C.c_void_p
    L.orc_scene_create.restype = vp
    L.orc_scene_destroy.argtypes = [vp]
    L.orc_scene_error.restype = C.c_char_p
    L.orc_scene_error.argtypes = [vp]
    L.orc_scene_add_obj.argtypes = [vp, C.c_char_p, C.c_char_p]
    L.orc_scene_add_arrays.argtypes = [vp, f32p, i32p, i32p, C.c_int, f32p, C.c_int, C.c_int]
    for n in ("n_tris", "n_mats", "n_lights", "n_objects"):
        getattr(L, "orc_scene_" + n).argtypes = [vp]
    L.orc_scene_get_tris.argtypes = [vp] + [vp] * 6
    L.orc_scene_get_mats.argtypes = [vp, f32p]
    L.orc_scene_light_size.argtypes = [vp, C.c_int]
    L.orc_scene_light_area.argtypes = [vp, C.c_int]
    L.orc_scene_light_area.restype = C.c_float
    L.orc_scene_light_tris.argtypes = [vp, C.c_int, i32p]
    L.orc_refbvh_build.argtypes = [vp, C.c_uint]
    L.orc_refbvh_root.argtypes = [vp]
    L.orc_refbvh_get.argtypes = [vp, vp, vp]
    L.orc_newbvh_build.argtypes = [vp, C.c_uint, C.c_int]
    L.orc_newbvh_get.argtypes = [vp, vp, vp, vp, vp]
    L.orc_trace.argtypes = [vp, C.c_int, C.c_int, f32p, C.c_int64, f32p, i32p, vp, C.c_int]
    L.orc_tri_test.argtypes = [vp, f32p, i32p, C.c_int64, f32p, vp]
    L.orc_primary_rays.argtypes = [f32p, f32p, C.c_float, C.c_int, C.c_int, f32p]
    L.orc_inverse_view_matrix.argtypes = [f32p, f32p, f32p, f32p]
    L.orc_render.argtypes = [vp, f32p, f32p, C.c_float, C.c_int, C.c_int, C.c_uint32, C.c_uint32, C.c_float,
                             C.c_int, C.c_uint32, C.c_int, C.c_int, i64p, vp, C.c_int]
    L.orc_wide8_build.argtypes = [vp, C.c_uint]
    L.orc_wide8_build2.argtypes = [vp, C.c_uint, C.c_int]
    L.orc_wide8_get.argtypes = [vp, vp, vp, vp, vp]
    L.orc_resolve.argtypes = [i64p, C.c_int, C.c_uint32, vp, vp]
    L.orc_philox.argtypes = [u32p, u32p, u32p]
    L.orc_u01.argtypes = [C.c_uint32]
    L.orc_u01.restype = C.c_float
    for fn, na in (("orc_det_log2", 1), ("orc_det_exp2", 1), ("orc_det_pow", 2)):
        getattr(L, fn).argtypes = [C.c_float] * na
        getattr(L, fn).restype = C.c_float
    L.orc_sincos_2pi.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    L.orc_sincos_rad.argtypes = [C.c_float, C.POINTER(C.c_float), C.POINTER(C.c_float)]
    _LIB = L
    return L


def _ptr(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


REF_NODE = np.dtype([("lc", "<i4"), ("rc", "<i4"), ("n", "<u4"), ("it", "<i4"), ("AA", "<f4", 3), ("BB", "<f4", 3)])
PAIR_NODE = np.dtype([("c0lox", "<f4"), ("c0hix", "<f4"), ("c0loy", "<f4"), ("c0hiy", "<f4"),
                      ("c1lox", "<f4"), ("c1hix", "<f4"), ("c1loy", "<f4"), ("c1hiy", "<f4"),
                      ("c0loz", "<f4"), ("c0hiz", "<f4"), ("c1loz", "<f4"), ("c1hiz", "<f4"),
                      ("c0", "<i4"), ("c1", "<i4"), ("n0", "<i4"), ("n1", "<i4")])
assert REF_NODE.itemsize == 40 and PAIR_NODE.itemsize == 64


def max_threads():
    return lib().orc_max_threads()


def inverse_view_matrix(eye, lookat, up):
    out = np.zeros(9, np.float32)
    lib().orc_inverse_view_matrix(np.asarray(eye, np.float32), np.asarray(lookat, np.float32),
                                  np.asarray(up, np.float32), out)
    return out


def primary_rays(eye, M, fovy_rad, width, height):
    rays = np.zeros((width * height, 8), np.float32)
    lib().orc_primary_rays(np.asarray(eye, np.float32), np.asarray(M, np.float32), fovy_rad, width, height, rays)
    return rays


class Scene:
    def __init__(self):
        self.L = lib()
        self.h = self.L.orc_scene_create()

    def __del__(self):
        if getattr(self, "h", None):
            self.L.orc_scene_destroy(self.h)
            self.h = None

    def add_obj(self, obj_path, mtl_dir):
        if self.L.orc_scene_add_obj(self.h, obj_path.encode(), mtl_dir.encode()) != 0:
            raise RuntimeError(self.L.orc_scene_error(self.h).decode())
        return self

    def add_arrays(self, verts, mat_id, obj_id, mats):
        verts = np.ascontiguousarray(verts, np.float32).reshape(-1, 9)
        mats = np.ascontiguousarray(mats, np.float32)          # rows of (kd, ke, ns) or (kd, ks, ke, ns)
        assert mats.ndim == 2 and mats.shape[1] in (7, 10)
        self.L.orc_scene_add_arrays(self.h, verts, np.ascontiguousarray(mat_id, np.int32),
                                    np.ascontiguousarray(obj_id, np.int32), verts.shape[0], mats, mats.shape[0], mats.shape[1])
        return self

    @property
    def n_tris(self):
        return self.L.orc_scene_n_tris(self.h)

    @property
    def n_lights(self):
        return self.L.orc_scene_n_lights(self.h)

    @property
    def n_objects(self):
        return self.L.orc_scene_n_objects(self.h)

    def tris(self):
        n = self.n_tris
        d = dict(verts=np.zeros((n, 9), np.float32), normal=np.zeros((n, 3), np.float32), area=np.zeros(n, np.float32),
                 area_of_obj=np.zeros(n, np.float32), mat=np.zeros(n, np.int32), obj=np.zeros(n, np.int32))
        self.L.orc_scene_get_tris(self.h, *[_ptr(d[k]) for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj")])
        return d

    def mats(self):
        out = np.zeros((self.L.orc_scene_n_mats(self.h), 9), np.float32)
        self.L.orc_scene_get_mats(self.h, out)
        return out

    def lights(self):
        res = []
        for li in range(self.n_lights):
            faces = np.zeros(self.L.orc_scene_light_size(self.h, li), np.int32)
            self.L.orc_scene_light_tris(self.h, li, faces)
            res.append((faces, self.L.orc_scene_light_area(self.h, li)))
        return res

    def build_ref_bvh(self, thresh_n):
        n = self.L.orc_refbvh_build(self.h, thresh_n)
        nodes = np.zeros(n, REF_NODE)
        order = np.zeros(self.n_tris, np.int32)
        self.L.orc_refbvh_get(self.h, _ptr(nodes), _ptr(order))
        return nodes, order, self.L.orc_refbvh_root(self.h)

    def build_new_bvh(self, thresh_n, builder=0):
        """64-byte pair nodes. builder: 0 = Karras radix tree (LBVH), 2 = PLOC topology."""
        n = self.L.orc_newbvh_build(self.h, thresh_n, builder)
        nodes = np.zeros(n, PAIR_NODE)
        order = np.zeros(self.n_tris, np.int32)
        last = np.zeros(self.n_tris, np.uint8)
        bounds = np.zeros(6, np.float32)
        self.L.orc_newbvh_get(self.h, _ptr(nodes), _ptr(order), _ptr(last), _ptr(bounds))
        return nodes, order, last, bounds

    def trace(self, rays, which=0, mode=0, threads=None, want_stats=False):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        n = rays.shape[0]
        t = np.zeros(n, np.float32)
        face = np.zeros(n, np.int32)
        stats = np.zeros(5, np.uint64)
        rc = self.L.orc_trace(self.h, which, mode, rays, n, t, face, _ptr(stats), threads or max_threads())
        if rc != 0:
            raise RuntimeError("orc_trace: BVH not built")
        if want_stats:
            return t, face, dict(zip(("inner", "boxes", "tris", "max_stack", "rays"), (int(x) for x in stats)))
        return t, face

    def tri_test(self, rays, faces):
        """Canonical triangle test of faces[k] against rays[k]: (t, inside)."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        faces = np.ascontiguousarray(faces, np.int32)
        t = np.zeros(len(faces), np.float32)
        inside = np.zeros(len(faces), np.uint8)
        self.L.orc_tri_test(self.h, rays, faces, len(faces), t, _ptr(inside))
        return t, inside.astype(bool)

    def check_any_hits(self, rays, t, face):
        """Any-hit results are 'a blocker or none' (which blocker is found first depends on warp scheduling):
        blocked status must equal this oracle's, a reported blocker must pass the canonical test with
        t > 1e-5 and tmax - t > 1e-5 (reference Render.cuh:19-27), and its t must be bit-exact."""
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        ot, of = self.trace(rays, which=0, mode=1)
        assert np.array_equal(face >= 0, of >= 0), "blocked status differs for %d rays" % int(((face >= 0) != (of >= 0)).sum())
        miss = face < 0
        assert np.all(t[miss] == np.float32(3.4028234663852886e38))
        tt, inside = self.tri_test(rays, face)
        b = ~miss
        eps = np.float32(0.00001)
        assert np.all(inside[b]) and np.array_equal(tt[b].view(np.uint32), t[b].view(np.uint32))
        assert np.all(t[b] > eps) and np.all(rays[b, 3] - t[b] > eps)
        return True

    def build_wide8(self, thresh_n, builder=1):
        """8-wide compressed BVH (80-byte nodes): returns nodes (n x 20 uint32), order, last, bounds.
        builder: 1 = collapse of the Karras tree (LBVH8), 3 = collapse of the PLOC tree (PLOC8)."""
        n = self.L.orc_wide8_build2(self.h, thresh_n, builder)
        nodes = np.zeros((n, 20), np.uint32)
        order = np.zeros(self.n_tris, np.int32)
        last = np.zeros(self.n_tris, np.uint8)
        bounds = np.zeros(6, np.float32)
        self.L.orc_wide8_get(self.h, _ptr(nodes), _ptr(order), _ptr(last), _ptr(bounds))
        return nodes, order, last, bounds

    def render(self, eye, M, fovy_rad, width, height, s_begin, s_end, p_rr, light_sample_n, seed=0, estimator=0,
               threads=None, accum=None, wide=False):
        if accum is None:
            accum = np.zeros(width * height * 3, np.int64)
        stats = np.zeros(12, np.uint64)
        rc = self.L.orc_render(self.h, np.asarray(eye, np.float32), np.asarray(M, np.float32), fovy_rad, width, height,
                               s_begin, s_end, p_rr, light_sample_n, seed, estimator, 1 if wide else 0, accum, _ptr(stats),
                               threads or max_threads())
        if rc != 0:
            raise RuntimeError("orc_render failed rc=%d" % rc)
        keys = ("samples", "extend_rays", "shadow_rays", "probe_rays", "closest_inner", "closest_tris", "closest_rays",
                "closest_max_stack", "any_inner", "any_tris", "any_rays", "any_max_stack")
        return accum, dict(zip(keys, (int(x) for x in stats)))


def resolve(accum, n_pixels, spp):
    lin = np.zeros(n_pixels * 3, np.float32)
    rgb = np.zeros(n_pixels * 3, np.uint8)
    lib().orc_resolve(np.ascontiguousarray(accum, np.int64), n_pixels, spp, _ptr(lin), _ptr(rgb))
    return lin, rgb


def philox(ctr, key):
    out = np.zeros(4, np.uint32)
    lib().orc_philox(np.asarray(ctr, np.uint32), np.asarray(key, np.uint32), out)
    return out


def sincos_2pi(u):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincos_2pi(u, C.byref(s), C.byref(c))
    return s.value, c.value


def sincos_rad(x):
    s, c = C.c_float(), C.c_float()
    lib().orc_sincos_rad(x, C.byref(s), C.byref(c))
    return s.value, c.value


def det_log2(x):
    return float(lib().orc_det_log2(float(x)))


def det_exp2(x):
    return float(lib().orc_det_exp2(float(x)))


def det_pow(x, y):
    return float(lib().orc_det_pow(float(x), float(y)))
