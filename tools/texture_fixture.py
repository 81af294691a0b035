"""Textured test scene for map_Kd (reference Loader.h:55-59,78-105; OBJLoader.h:184-193): OBJ + MTL + texture files
written from a seed. The PNG encoder below exists only to produce test inputs of every kind the product's reader
accepts (colour types 0/2/3/4/6, 1-16 bits, all five row filters, tRNS).

The reference's own result for a textured triangle is UNDEFINED: `auto kd_1 = Eigen::Vector3f(...) / 255.;`
(Loader.h:88,94,100) keeps an Eigen expression that refers to a destroyed temporary. Built with -O2 it yields
~1e-12 (black); built with -O0 all three expressions alias the last temporary and Kd becomes the THIRD corner's texel.
The product implements the evident intent (mean of the three texels). What can still be pinned on the real reference
code is everything before the mean - stb_image's decode, the (x, y) swap, the uv -> texel arithmetic, the /255 - by a
scene whose triangles carry the same uv on all three corners: there intent and the -O0 build agree exactly.

    python tools/texture_fixture.py        regenerates tests/golden/map_kd.npz from the REAL reference loader built
                                           with -O0 (oracle/ref_harness/build.sh O0; needs /root/reference) on the
                                           same-uv scene - run in the build container.
"""
import os
import struct
import sys
import zlib

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _chunk(kind, data):
    return struct.pack(">I", len(data)) + kind + data + struct.pack(">I", zlib.crc32(kind + data) & 0xffffffff)


def _paeth(a, b, c):
    p = a + b - c
    pa, pb, pc = abs(p - a), abs(p - b), abs(p - c)
    return a if (pa <= pb and pa <= pc) else (b if pb <= pc else c)


def write_png(path, samples, ctype, depth=8, palette=None, trns=None, idat_split=1):
    """samples: (H, W, C) integer array of raw sample values (palette indices for ctype 3). Rows cycle through the
    five filter types, so the reader's unfiltering is exercised."""
    samples = np.asarray(samples)
    H, W, C = samples.shape
    assert C == {0: 1, 2: 3, 3: 1, 4: 2, 6: 4}[ctype]
    rows = []
    for y in range(H):
        flat = samples[y].reshape(-1)
        if depth == 16:
            b = bytearray()
            for v in flat:
                b += struct.pack(">H", int(v))
        elif depth == 8:
            b = bytearray(int(v) for v in flat)
        else:
            per = 8 // depth
            b = bytearray((len(flat) + per - 1) // per)
            for i, v in enumerate(flat):
                b[i // per] |= (int(v) & ((1 << depth) - 1)) << (8 - depth - (i % per) * depth)
        rows.append(bytes(b))
    bpp = max(1, C * depth // 8)
    raw = bytearray()
    prev = bytes(len(rows[0]))
    for y, cur in enumerate(rows):
        ft = y % 5
        out = bytearray(len(cur))
        for i in range(len(cur)):
            a = cur[i - bpp] if i >= bpp else 0
            b = prev[i]
            c = prev[i - bpp] if i >= bpp else 0
            pred = (0, a, b, (a + b) >> 1, _paeth(a, b, c))[ft]
            out[i] = (cur[i] - pred) & 0xff
        raw.append(ft)
        raw += out
        prev = cur
    z = zlib.compress(bytes(raw), 6)
    png = b"\x89PNG\r\n\x1a\n" + _chunk(b"IHDR", struct.pack(">IIBBBBB", W, H, depth, ctype, 0, 0, 0))
    if palette is not None:
        png += _chunk(b"PLTE", bytes(int(v) for v in np.asarray(palette).reshape(-1)))
    if trns is not None:
        png += _chunk(b"tRNS", bytes(trns))
    step = (len(z) + idat_split - 1) // idat_split
    for k in range(0, len(z), step):
        png += _chunk(b"IDAT", z[k:k + step])
    png += _chunk(b"IEND", b"")
    with open(path, "wb") as f:
        f.write(png)


def write_scene(d, seed=7, n=12, same_uv=False):
    """Writes textured.obj / textured.mtl / textures into directory d. Returns (OBJ path, info) with
    info["pixels"][file] = the decoded image (H, W, C) uint8 as stb_image reports it, info["groups"] = [(material,
    file or None, plain Kd)], info["uv"] = (T, 3, 2) float32 corner uvs in face order (NaN rows for the lamp).
    One vt per vertex (the reference indexes vt by the vertex index); uv inside and outside [0, 1), negative ones too.
    same_uv: every triangle has its own three vertices, all with the same uv (see the module text)."""
    rng = np.random.default_rng(seed)
    os.makedirs(d, exist_ok=True)
    px = {}
    a = rng.integers(0, 256, (5, 7, 3)); write_png(os.path.join(d, "rgb.png"), a, 2, 8, idat_split=3); px["rgb.png"] = a   # non-square
    a = rng.integers(0, 256, (8, 8, 4)); write_png(os.path.join(d, "rgba.png"), a, 6, 8); px["rgba.png"] = a
    a = rng.integers(0, 65536, (6, 6, 3)); write_png(os.path.join(d, "rgb16.png"), a, 2, 16); px["rgb16.png"] = a >> 8
    a, pal = rng.integers(0, 16, (9, 6, 1)), rng.integers(0, 256, (16, 3))
    write_png(os.path.join(d, "pal4.png"), a, 3, 4, palette=pal); px["pal4.png"] = pal[a[:, :, 0]]
    a, pal, tr = rng.integers(0, 8, (4, 4, 1)), rng.integers(0, 256, (8, 3)), np.array([0, 128, 255, 255, 255, 255, 255, 255])
    write_png(os.path.join(d, "palt.png"), a, 3, 8, palette=pal, trns=[0, 128, 255])
    px["palt.png"] = np.concatenate([pal[a[:, :, 0]], tr[a[:, :, 0]][:, :, None]], axis=2)
    a = rng.integers(0, 256, (10, 10, 2)); write_png(os.path.join(d, "ga.png"), a, 4, 8); px["ga.png"] = a
    a = rng.integers(0, 4, (6, 5, 1)); write_png(os.path.join(d, "g2.png"), a, 0, 2); px["g2.png"] = a * 0x55
    a = rng.integers(0, 256, (4, 5, 3))
    with open(os.path.join(d, "p6.ppm"), "wb") as f:
        f.write(b"P6\n# comment\n5 4\n255\n" + bytes(int(v) for v in a.reshape(-1)))
    px["p6.ppm"] = a
    px = {k: np.ascontiguousarray(v, np.uint8) for k, v in px.items()}
    groups = [("t_rgb", "rgb.png"), ("t_rgba", "rgba.png"), ("t_rgb16", "rgb16.png"), ("t_pal4", "pal4.png"), ("t_palt", "palt.png"),
              ("t_ga", "ga.png"), ("t_g2", "g2.png"), ("t_ppm", "p6.ppm"), ("t_missing", "no_such_file.png"), ("plain", None)]
    with open(os.path.join(d, "textured.mtl"), "w") as f:
        for k, (name, tex) in enumerate(groups):
            f.write("newmtl %s\nKd %.3f %.3f %.3f\nKe 0 0 0\nNs 1\n" % (name, 0.1 + 0.05 * k, 0.9 - 0.05 * k, 0.5))
            if tex:
                f.write("map_Kd %s\n" % tex)
            f.write("\n")
        f.write("newmtl lamp\nKd 0 0 0\nKe 10 10 10\nNs 1\n")
    uvs, kds = [], []
    with open(os.path.join(d, "textured.obj"), "w") as f:
        f.write("mtllib textured.mtl\n")
        nv = 0
        for g, (name, tex) in enumerate(groups):
            plain = (np.float32("%.3f" % (0.1 + 0.05 * g)), np.float32("%.3f" % (0.9 - 0.05 * g)), np.float32(0.5))
            # a strip of (n+1) x 2 vertices per group, z = group index
            uv = rng.uniform(-1.5, 2.5, ((n + 1) * 2, 2))
            if name in ("t_ga", "t_g2"):
                uv = rng.uniform(0.0, 0.8, ((n + 1) * 2, 2))       # 1/2-channel texels read the neighbours' bytes: stay off the last texel
            else:
                uv[0] = (0.0, 1.0)                                  # exact integers: frac01 gives 0
                uv[1] = (-0.25, 0.999999)
            uv = np.array([["%.7g" % x for x in r] for r in uv]).astype(np.float32)       # what the file says
            P = [(i, j + 1, g) for i in range(n + 1) for j in range(2)]
            faces = []
            for i in range(n):
                a, b, c, e = 2 * i, 2 * i + 1, 2 * i + 2, 2 * i + 3
                faces += [(a, c, b), (b, c, e)]
            if same_uv:                                            # own vertices per face, uv of the face's third corner on all of them
                P2, uv2, faces2 = [], [], []
                for (a, b, c) in faces:
                    k = len(P2)
                    P2 += [P[a], P[b], P[c]]
                    uv2 += [uv[c]] * 3
                    faces2.append((k, k + 1, k + 2))
                P, uv, faces = P2, np.array(uv2, np.float32), faces2
            for q in P:
                f.write("v %g %g %g\n" % q)
            for a in uv:
                f.write("vt %.9g %.9g\n" % (a[0], a[1]))
            for a in P:
                f.write("vn 0 0 1\n")
            f.write("usemtl %s\n" % name)
            for (a, b, c) in faces:
                a, b, c = a + nv + 1, b + nv + 1, c + nv + 1
                f.write("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c))
                uvs.append([uv[a - nv - 1], uv[b - nv - 1], uv[c - nv - 1]])
                kds.append(plain)
            nv += len(P)
        for v in ((0, 5, 0), (4, 5, 0), (4, 5, 4)):
            f.write("v %g %g %g\nvt 0 0\nvn 0 -1 0\n" % v)
        f.write("usemtl lamp\nf %d/%d/%d %d/%d/%d %d/%d/%d\n" % ((nv + 1,) * 3 + (nv + 2,) * 3 + (nv + 3,) * 3))
        uvs.append([[np.nan, np.nan]] * 3)
        kds.append((0.0, 0.0, 0.0))
    per_face_tex = []
    for name, tex in groups:
        per_face_tex += [tex if (tex in px) else None] * (2 * n)
    per_face_tex.append(None)
    info = dict(pixels=px, groups=groups, uv=np.array(uvs, np.float32), plain_kd=np.array(kds, np.float32), face_texture=per_face_tex)
    return os.path.join(d, "textured.obj"), info


def reference_kd(obj, d, so=None):
    """Per-triangle record (23 floats) from the REAL reference loader, scene order."""
    sys.path.insert(0, ROOT)
    import ctypes as C
    from oracle import ref
    if so:
        ref.SO = so
    R = ref._lib()
    R.ref_n_tris.restype = C.c_int
    with ref.quiet_stdout():
        h = R.ref_host_load(obj.encode(), (d.rstrip("/") + "/").encode(), 64, 64, 2)
    n = R.ref_n_tris(h)
    t = np.zeros((n, 23), np.float32)
    R.ref_get_tris(h, 0, t.ctypes.data_as(C.c_void_p))
    return t


if __name__ == "__main__":
    import subprocess
    import tempfile
    subprocess.check_call(["bash", os.path.join(ROOT, "oracle", "ref_harness", "build.sh"), "O0"])
    tmp = tempfile.mkdtemp()
    obj, info = write_scene(tmp, same_uv=True)
    t = reference_kd(obj, tmp, os.path.join(ROOT, "oracle", "_ref", "libref_O0.so"))
    out = os.path.join(ROOT, "tests", "golden", "map_kd.npz")
    np.savez_compressed(out, verts=t[:, 0:9], kd=t[:, 14:17], ke=t[:, 17:20])
    print("wrote", out, t.shape, "distinct kd:", len(np.unique(t[:, 14:17], axis=0)))
