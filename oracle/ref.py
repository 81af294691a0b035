"""ORACLE — TEST INFRASTRUCTURE ONLY.

ctypes binding of oracle/_ref/libref.so: the REFERENCE's own headers (from /root/reference/include, compiled in
place by oracle/ref_harness/build.sh) behind a small headless harness. Host functions (OBJ load, BVH build)
run anywhere; device functions (view_render_kernel, DeviceBVH::intersect) need a GPU. Used by tests/, by
bench.py --impl reference and by tools/make_golden.py — never by the product.
"""
import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
SO = os.path.join(_HERE, "_ref", "libref.so")


def available():
    return os.path.exists(SO)


def _lib():
    R = C.CDLL(SO)
    vp = C.c_void_p
    R.ref_host_load.restype = vp
    R.ref_host_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint]
    R.ref_host_times.argtypes = [vp, vp]
    for f in ("ref_n_tris", "ref_n_nodes", "ref_root", "ref_n_lights", "ref_device_init", "ref_free"):
        getattr(R, f).argtypes = [vp]
    R.ref_get_tris.argtypes = [vp, C.c_int, vp]
    R.ref_get_nodes.argtypes = [vp, vp]
    R.ref_get_light.argtypes = [vp, C.c_int, vp, vp]
    R.ref_inverse_view.argtypes = [vp] * 4
    R.ref_render.argtypes = [vp, vp, vp, C.c_float, C.c_uint, C.c_float, C.c_int, vp, C.POINTER(C.c_float), C.POINTER(C.c_double)]
    R.ref_trace.argtypes = [vp, vp, C.c_longlong, vp, C.POINTER(C.c_float)]
    return R


class quiet_stdout:
    """The reference printf()s from its loaders and constructors; keep the bench's stdout to one JSON line."""

    def __enter__(self):
        self.saved = os.dup(1)
        self.null = os.open(os.devnull, os.O_WRONLY)
        os.dup2(self.null, 1)

    def __exit__(self, *a):
        os.dup2(self.saved, 1)
        os.close(self.null)
        os.close(self.saved)


class RefScene:
    """Loader + Scene + BVH (host), optionally DeviceBVH/DeviceLights/stacks (device) of the reference."""

    def __init__(self, obj_path, mtl_dir, width, height, thresh_n):
        self.R = _lib()
        self.width, self.height = width, height
        if not mtl_dir.endswith("/"):
            mtl_dir += "/"
        with quiet_stdout():
            self.h = self.R.ref_host_load(obj_path.encode(), mtl_dir.encode(), width, height, thresh_n)
        self.device_ready = False

    def host_times_ms(self):
        ms = np.zeros(3)
        self.R.ref_host_times(self.h, ms.ctypes.data_as(C.c_void_p))
        return dict(obj_parse=float(ms[0]), load_object=float(ms[1]), bvh_build=float(ms[2]))

    def device_init(self):
        with quiet_stdout():
            rc = self.R.ref_device_init(self.h)
        self.device_ready = rc == 0
        return rc

    def inverse_view(self, eye, lookat, up):
        e, l, u = (np.ascontiguousarray(x, np.float32) for x in (eye, lookat, up))
        out = np.zeros(9, np.float32)
        self.R.ref_inverse_view(*(a.ctypes.data_as(C.c_void_p) for a in (e, l, u, out)))
        return out

    def render(self, eye, M, fovy_rad, spp, p_rr, lsn, frame=None):
        """view_render_kernel; returns (rgb8 frame, kernel_ms, wall_ms)."""
        if frame is None:
            frame = np.zeros((self.height, self.width, 3), np.uint8)
        e = np.ascontiguousarray(eye, np.float32)
        m = np.ascontiguousarray(M, np.float32)
        kms, wms = C.c_float(), C.c_double()
        rc = self.R.ref_render(self.h, e.ctypes.data_as(C.c_void_p), m.ctypes.data_as(C.c_void_p), float(fovy_rad), int(spp), float(p_rr),
                               int(lsn), frame.ctypes.data_as(C.c_void_p), C.byref(kms), C.byref(wms))
        if rc != 0:
            raise RuntimeError("reference render failed rc=%d" % rc)
        return frame, kms.value, wms.value

    def trace(self, rays):
        rays = np.ascontiguousarray(rays, np.float32).reshape(-1, 8)
        t = np.zeros(len(rays), np.float32)
        ms = C.c_float()
        rc = self.R.ref_trace(self.h, rays.ctypes.data_as(C.c_void_p), len(rays), t.ctypes.data_as(C.c_void_p), C.byref(ms))
        if rc != 0:
            raise RuntimeError("reference trace failed rc=%d" % rc)
        return t, ms.value

    def close(self):
        if getattr(self, "h", None):
            self.R.ref_free(self.h)
            self.h = None
