"""Generates tests/golden/* in THIS container (needs /root/reference and oracle/_ref/libref.so).

  * scene fixtures (tests/golden/scenes/*.npz.xz) from the reference's OBJ/MTL/config files;
  * golden.json: digests produced by the REAL reference host code (OBJ loader, Triangle, Object, BVH.h)
    through oracle/_ref/libref.so, plus oracle-produced vectors pinned for regression;
  * <scene>_hits_200x150.npz: primary-ray hit ids / t bits (reference BVH + reference traversal rule,
    canonical triangle arithmetic).
Run:  python tools/make_golden.py
"""
import ctypes as C
import hashlib
import json
import math
import os
import sys
import tempfile

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import orc
from tools import scene_fixture as sf

REF = "/root/reference"


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


def main():
    R = C.CDLL(os.path.join(ROOT, "oracle", "_ref", "libref.so"))
    R.ref_host_load.restype = C.c_void_p
    R.ref_host_load.argtypes = [C.c_char_p, C.c_char_p, C.c_uint, C.c_uint, C.c_uint]
    for f in ("ref_n_tris", "ref_n_nodes", "ref_root", "ref_n_lights"):
        getattr(R, f).argtypes = [C.c_void_p]
    R.ref_get_tris.argtypes = [C.c_void_p, C.c_int, C.c_void_p]
    R.ref_get_nodes.argtypes = [C.c_void_p, C.c_void_p]
    R.ref_get_light.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
    golden = {}
    for name in ("cornell-box", "veach-mis"):
        d = "%s/scenes/%s/" % (REF, name)
        sf.pack(d + name + ".obj", d, d + "config.json", sf.fixture(name))
        with open(d + "config.json") as f:
            cfg = json.load(f)
        thresh = cfg["bvh_thresh_n"]
        # --- the real reference host path
        h = R.ref_host_load((d + name + ".obj").encode(), d.encode(), cfg["width"], cfg["height"], thresh)
        n, nn = R.ref_n_tris(h), R.ref_n_nodes(h)
        t0 = np.zeros((n, 23), np.float32); R.ref_get_tris(h, 0, t0.ctypes.data_as(C.c_void_p))
        t1 = np.zeros((n, 23), np.float32); R.ref_get_tris(h, 1, t1.ctypes.data_as(C.c_void_p))
        nodes = np.zeros(nn, orc.REF_NODE); R.ref_get_nodes(h, nodes.ctypes.data_as(C.c_void_p))
        areas, sizes = [], []
        for li in range(R.ref_n_lights(h)):
            nt, ar = C.c_int(), C.c_float()
            R.ref_get_light(h, li, C.byref(nt), C.byref(ar))
            areas.append(float(np.float32(ar.value))); sizes.append(nt.value)
        cam = dict(eye=[cfg["eye_pos"][k] for k in "xyz"], lookat=[cfg["lookat"][k] for k in "xyz"], up=[cfg["up"][k] for k in "xyz"],
                   fov_y=cfg["fov_y"], P_RR=cfg["P_RR"], light_sample_n=cfg["light_sample_n"], spp=cfg["spp"])
        g = dict(n_tris=n, n_lights=len(areas), light_areas=areas, light_sizes=sizes, thresh_n=thresh, camera=cam,
                 ref_verts_sha256=sha(t0[:, 0:9]), ref_normal_sha256=sha(t0[:, 9:12]), ref_area_sha256=sha(t0[:, 12]),
                 ref_n_nodes=nn, ref_root=R.ref_root(h), ref_nodes_sha256=sha(nodes), ref_sorted_verts_sha256=sha(t1[:, 0:9]))
        # --- oracle vectors on the unpacked fixture (what the tests load)
        tmp = tempfile.mkdtemp()
        sf.unpack(sf.fixture(name), tmp)
        S = orc.Scene().add_obj(os.path.join(tmp, name + ".obj"), tmp)
        S.build_ref_bvh(thresh); S.build_new_bvh(thresh)
        M = orc.inverse_view_matrix(cam["eye"], cam["lookat"], cam["up"])
        fov = float(np.float32(np.float32(cam["fov_y"]) * np.float32(math.pi) / np.float32(180)))
        rays = orc.primary_rays(cam["eye"], M, fov, 200, 150)
        t, face, st = S.trace(rays, which=1, want_stats=True)
        np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + "_hits_200x150.npz"), face=face, t_bits=t.view(np.uint32))
        g["hit_fraction_200x150"] = float((face >= 0).mean())
        g["ref_pops_per_ray_200x150"] = st["inner"] / st["rays"]
        acc, _ = S.render(cam["eye"], M, fov, 64, 48, 0, 4, cam["P_RR"], cam["light_sample_n"], seed=0)
        g["oracle_accum_64x48_spp4_sha256"] = sha(acc)
        golden[name] = g
        print(name, {k: v for k, v in g.items() if not isinstance(v, (dict, list))})
    with open(os.path.join(ROOT, "tests", "golden", "golden.json"), "w") as f:
        json.dump(golden, f, indent=1)


if __name__ == "__main__":
    main()
