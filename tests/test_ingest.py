"""CPU tests of the product's host side and of the C-ABI surface (no GPU compute calls)."""
import ctypes as C
import json
import os
import re
import struct
import zlib

import numpy as np
import pytest

from conftest import ROOT, box_scene


def test_cabi_exports_every_declared_symbol(crt):
    hdr = open(os.path.join(ROOT, "include", "crt.h")).read()
    names = sorted(set(re.findall(r"\b(crt_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 35
    L = C.CDLL(crt.lib_path())
    for n in names:
        assert hasattr(L, n), "libcrt.so does not export " + n
    assert L.crt_abi_version() == 2          # 2: crt_group (several GPUs behind one handle)


def test_no_gpu_means_loud_failure_not_fallback(crt):
    if crt.device_count() > 0:
        pytest.skip("a GPU is present")
    s = crt.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [0], [0], [[.5, .5, .5, 0, 0, 0, 1]])
    with pytest.raises(crt.CrtError) as e:
        s.set_BVH(2)
    assert e.value.code == -3 and "no CPU fallback" in str(e.value)
    with pytest.raises(crt.CrtError) as e:
        crt.Render(s, 8, 8)
    assert e.value.code == -5                      # render before build_bvh: state error
    with pytest.raises(crt.CrtError):
        s.trace_rays(np.zeros((1, 8), np.float32))


def test_product_never_imports_the_oracle():
    """The oracle is test infrastructure: no include, import, link or dlopen of it from the product."""
    pkg = os.path.join(ROOT, "cudaraytracing_b200")
    bad = re.compile(r'#\s*include\s*[<"][^>"]*(orc_|oracle)|^\s*(from|import)\s+oracle|liborc|-lorc|CDLL\([^)]*oracle', re.M)
    n = 0
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")) or f == "Makefile":
                txt = open(os.path.join(dp, f), errors="replace").read()
                assert not bad.search(txt), f
                n += 1
    assert n >= 10


@pytest.mark.parametrize("name", ["cornell-box", "veach-mis"])
def test_obj_ingest_matches_oracle_restatement(crt, orc, scene_files, name):
    f = scene_files[name]
    a = crt.Scene().add_obj(f["obj"], f["dir"])
    b = orc.Scene().add_obj(f["obj"], f["dir"])
    ta, tb = a.tris(), b.tris()
    for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj"):
        assert np.array_equal(ta[k].view(np.uint32), tb[k].view(np.uint32)), k
    assert np.array_equal(a.mats(), b.mats())
    la, lb = a.lights(), b.lights()
    assert len(la) == len(lb)
    for (fa, aa), (fb, ab) in zip(la, lb):
        assert np.array_equal(fa, fb) and np.float32(aa) == np.float32(ab)


def test_obj_loader_quirks(crt, orc, tmp_path):
    """Reference loader behaviour: faces before the first usemtl are dropped, polygons keep their first
    three corners, v/vt/vn and v//vn and bare indices parse, unknown materials stay black, Ks is ignored by
    compat, Ns>1 switches the mode, CRLF line ends are fine (OBJLoader.h:61-203, Loader.h:40-124)."""
    obj = ("mtllib m.mtl\r\nv 0 0 0\nv 1 0 0\nv 0 1 0\nv 0 0 1\nvn 0 0 1\nvt 0 0\n"
           "f 1 2 3\n"                      # dropped: no usemtl yet
           "usemtl a\nf 1/1/1 2/1/1 3/1/1 4/1/1\nf 1//1 2//1 4//1\n"
           "usemtl glow\nf 2 3 4\n"
           "usemtl missing\nf 1 3 4\n"
           "usemtl a\nf 3 2 1\n")
    mtl = "newmtl a\nKd 0.25 0.5 0.75\nKs 1 1 1\nNs 50\nnewmtl glow\nKe 2 0 0\nKd 0 0 0\nNs 1\n"
    (tmp_path / "s.obj").write_text(obj)
    (tmp_path / "m.mtl").write_text(mtl)
    a = crt.Scene().add_obj(str(tmp_path / "s.obj"), str(tmp_path))
    b = orc.Scene().add_obj(str(tmp_path / "s.obj"), str(tmp_path))
    assert a.counts()["n_tris"] == 5 and b.n_tris == 5
    ta, tb = a.tris(), b.tris()
    for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj"):
        assert np.array_equal(ta[k].view(np.uint32), tb[k].view(np.uint32)), k
    m = a.mats()
    assert np.array_equal(m, b.mats())
    assert m[0][8] == 1 and m[0][6] == 50 and list(m[0][:3]) == [0.25, 0.5, 0.75]       # SPECULAR from Ns
    assert m[1][7] == 1 and m[2][7] == 0 and list(m[2][:6]) == [0] * 6                  # emissive; unknown = black
    assert ta["obj"].tolist() == [0, 0, 1, 2, 3]
    assert [len(f) for f, _ in a.lights()] == [1]
    assert np.array_equal(ta["verts"][0], np.array([0, 0, 0, 1, 0, 0, 0, 1, 0], np.float32))   # first 3 corners of the quad


def test_ingest_errors(crt, tmp_path):
    with pytest.raises(crt.CrtError) as e:
        crt.Scene().add_obj(str(tmp_path / "nope.obj"), str(tmp_path))
    assert e.value.code == -2 and "Unable to open OBJ file" in str(e.value)
    (tmp_path / "a.obj").write_text("mtllib gone.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl x\nf 1 2 3\n")
    with pytest.raises(crt.CrtError) as e:
        crt.Scene().add_obj(str(tmp_path / "a.obj"), str(tmp_path))
    assert "Unable to open MTL file" in str(e.value)
    (tmp_path / "b.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nusemtl x\nf 1 2 9\n")
    (tmp_path / "m.mtl").write_text("newmtl x\nKd 1 1 1\n")
    with pytest.raises(crt.CrtError) as e:
        crt.Scene().add_obj(str(tmp_path / "b.obj"), str(tmp_path))
    assert "malformed face" in str(e.value)
    with pytest.raises(crt.CrtError):
        crt.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, float("nan"), 0]], [0], [0], [[.5, .5, .5, 0, 0, 0, 1]])
    with pytest.raises(crt.CrtError):
        crt.Scene().add_triangles([[0, 0, 0, 1, 0, 0, 0, 1, 0]], [3], [0], [[.5, .5, .5, 0, 0, 0, 1]])


def test_add_triangles_matches_oracle(crt, orc):
    verts, mat, obj, mats = box_scene(np.random.default_rng(0), 50)
    a = crt.Scene().add_triangles(verts, mat, obj, mats)
    b = orc.Scene().add_arrays(verts, mat, obj, mats)
    ta, tb = a.tris(), b.tris()
    for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj"):
        assert np.array_equal(ta[k].view(np.uint32), tb[k].view(np.uint32)), k
    assert np.array_equal(a.mats(), b.mats())
    assert len(a.lights()) == len(b.lights()) == 1


def test_config_loader(crt, scene_files, tmp_path):
    cfg = crt.load_config(scene_files["cornell-box"]["cfg_path"])
    assert (cfg.width, cfg.height, cfg.spp, cfg.light_sample_n, cfg.bvh_thresh_n) == (800, 600, 2, 2, 2)
    assert cfg.P_RR == np.float32(0.6) and cfg.fov_y == np.float32(39.3077)
    assert list(cfg.eye_pos) == [278.0, 273.0, -800.0] and cfg.seed == 0 and cfg.estimator == 0
    v = crt.load_config(scene_files["veach-mis"]["cfg_path"])
    assert v.eye_pos[2] == np.float32(1.23612e-06) and v.spp == 4
    bad = tmp_path / "bad.json"
    bad.write_text('{"OBJ_paths": [], "width": 8}')
    with pytest.raises(crt.CrtError) as e:
        crt.load_config(str(bad))
    assert "missing" in str(e.value)
    bad.write_text("{ not json")
    with pytest.raises(crt.CrtError):
        crt.load_config(str(bad))
    ext = json.load(open(scene_files["veach-mis"]["cfg_path"]))
    ext.update(seed=7, estimator="mis")
    (tmp_path / "ext.json").write_text(json.dumps(ext))
    c2 = crt.load_config(str(tmp_path / "ext.json"))
    assert c2.seed == 7 and c2.estimator == 1


def test_camera_matrix_matches_oracle(crt, orc):
    for eye, look, up in ([[278, 273, -800], [278, 273, -799], [0, 1, 0]], [[28.2792, 5.2, 1.23612e-06], [0, 2.8, 0], [0, 1, 0]],
                          [[1, 2, 3], [-4, 0.5, 9], [0.1, 1, 0.2]]):
        assert np.array_equal(crt.inverse_view_matrix(eye, look, up), orc.inverse_view_matrix(eye, look, up))


def test_png_writer_round_trip(crt, tmp_path):
    rng = np.random.default_rng(5)
    img = rng.integers(0, 256, (37, 53, 3), dtype=np.uint8)
    p = str(tmp_path / "x.png")
    crt.write_png(p, img)
    data = open(p, "rb").read()
    assert data[:8] == b"\x89PNG\r\n\x1a\n"
    pos, idat, ihdr = 8, b"", None
    while pos < len(data):
        n, typ = struct.unpack(">I4s", data[pos:pos + 8])
        body = data[pos + 8:pos + 8 + n]
        assert struct.unpack(">I", data[pos + 8 + n:pos + 12 + n])[0] == (zlib.crc32(typ + body) & 0xffffffff)
        if typ == b"IHDR": ihdr = struct.unpack(">IIBBBBB", body)
        if typ == b"IDAT": idat += body
        pos += 12 + n
    assert ihdr == (53, 37, 8, 2, 0, 0, 0)
    raw = zlib.decompress(idat)
    rows = np.frombuffer(raw, np.uint8).reshape(37, 1 + 53 * 3)
    assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(37, 53, 3), img)


def test_cli_reports_errors(crt, tmp_path):
    import subprocess
    exe = os.path.join(ROOT, "cudaraytracing_b200", "crt")
    r = subprocess.run([exe, "--config", str(tmp_path / "missing.json")], capture_output=True, text=True)
    assert r.returncode == 1 and "cannot open" in r.stderr


def test_shard_work_partitions_exactly(crt):
    from cudaraytracing_b200.distributed import shard_work
    for npix, spp in ((800 * 600, 2), (800 * 600, 4), (3840 * 2160, 1024), (7, 3), (5, 1)):
        for world in (1, 2, 3, 4, 8):
            parts = [shard_work(npix, spp, r, world) for r in range(world)]
            assert parts[0][0] == 0 and parts[-1][1] == npix * spp
            assert all(parts[k][1] == parts[k + 1][0] for k in range(world - 1))
            if spp >= world:
                assert all(b % npix == 0 and e % npix == 0 for b, e in parts)


# ---- map_Kd textures (reference Loader.h:55-59,78-105) ----
def _kd_of(scene):
    tr, m = scene.tris(), scene.mats()
    return tr, m[tr["mat"]][:, 0:3]


def test_map_kd_equals_the_statement(crt, tmp_path):
    """Every kind of texture file the reader accepts (PNG colour types 0/2/3/4/6, 2/4/8/16 bits, five row filters, tRNS,
    split IDAT; binary PPM), uv inside / outside [0, 1) and negative: per-triangle Kd = the numpy statement of the
    reference's rule, bit for bit. A texture file that does not exist leaves the plain Kd (stbi_load returns null)."""
    from oracle import orc_texture as ot
    from tools import texture_fixture as tf
    obj, info = tf.write_scene(str(tmp_path))
    S = crt.Scene().add_obj(obj, str(tmp_path))
    tr, kd = _kd_of(S)
    assert len(kd) == len(info["uv"])
    want = info["plain_kd"].copy()
    for name, px in info["pixels"].items():
        rows = [i for i, t in enumerate(info["face_texture"]) if t == name]
        want[rows] = ot.kd_from_texture(px, info["uv"][rows])
    assert np.array_equal(kd.view(np.uint32), want.view(np.uint32))
    textured = np.array([t is not None for t in info["face_texture"]])
    assert len(np.unique(kd[textured], axis=0)) > 50                     # the textures do vary over the strips
    assert S.counts()["n_lights"] == 1 and S.mats().shape[0] > 20         # one material per distinct texel mean


def test_map_kd_equals_the_real_reference_on_the_pinnable_scene(crt, tmp_path):
    """tests/golden/map_kd.npz comes from the REAL reference loader (stb_image decode, swapped (x, y), uv -> texel
    arithmetic, /255), -O0 build, on the scene whose triangles carry one uv on all three corners - the only case in
    which the reference's own mean is not undefined behaviour in effect (tools/texture_fixture.py)."""
    from tools import texture_fixture as tf
    g = np.load(os.path.join(ROOT, "tests", "golden", "map_kd.npz"))
    obj, info = tf.write_scene(str(tmp_path), same_uv=True)
    tr, kd = _kd_of(crt.Scene().add_obj(obj, str(tmp_path)))
    assert np.array_equal(tr["verts"].view(np.uint32), g["verts"].view(np.uint32))
    assert np.array_equal(kd.view(np.uint32), g["kd"].view(np.uint32))
    assert len(np.unique(g["kd"], axis=0)) > 50


def test_map_kd_errors(crt, tmp_path):
    from tools import texture_fixture as tf
    obj, _ = tf.write_scene(str(tmp_path))
    with open(os.path.join(str(tmp_path), "rgb.png"), "wb") as f:        # exists, but is not an image we can decode: loud
        f.write(b"\xff\xd8\xff\xe0 this is not a PNG")
    with pytest.raises(crt.CrtError):
        crt.Scene().add_obj(obj, str(tmp_path))
    obj, _ = tf.write_scene(str(tmp_path))
    with open(os.path.join(str(tmp_path), "rgb.png"), "r+b") as f:       # truncated file
        f.truncate(60)
    with pytest.raises(crt.CrtError):
        crt.Scene().add_obj(obj, str(tmp_path))
    # fewer vt than vertices: the reference would index out of bounds (Loader.h:81-83)
    with open(os.path.join(str(tmp_path), "few.obj"), "w") as f:
        f.write("mtllib textured.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0.5 0.5\nusemtl t_rgba\nf 1 2 3\n")
    with pytest.raises(crt.CrtError):
        crt.Scene().add_obj(os.path.join(str(tmp_path), "few.obj"), str(tmp_path))


def _big_obj(path, n_quads=12000, bad_at=None, forward_ref_at=None, reference_forms_only=False, out_of_range_at=None):
    """An OBJ large enough to be read in several chunks: relative and absolute indices, faces before the first usemtl,
    several usemtl, two mtllib lines (the last one wins), v/vt/vn forms. Returns the number of lines."""
    rng = np.random.default_rng(11)
    lines = ["# generated", "mtllib wrong.mtl", "v 0 0 0", "v 1 0 0", "v 0 1 0", "f 1 2 3", "mtllib big.mtl"]   # dropped face
    for q in range(n_quads):
        x, y, z = rng.uniform(-50, 50, 3)
        lines += ["v %.6f %.6f %.6f" % (x, y, z), "v %.6f %.6f %.6f" % (x + 1, y, z), "v %.6f %.6f %.6f" % (x + 1, y + 1, z + 0.25),
                  "v %.6f %.6f %.6f" % (x, y + 1, z), "vt 0.5 0.5"]
        if q % 3000 == 0:
            lines.append("usemtl m%d" % (q // 3000 % 3))
        if reference_forms_only:            # what the reference's stoull-based reader accepts: positive v or v/vt/vn
            base = 3 + 4 * q
            lines += ["f %d/1/1 %d/1/1 %d/1/1" % (base + 1, base + 2, base + 3), "f %d %d %d %d" % (base + 1, base + 3, base + 4, base + 2)]
        elif q % 2:
            lines += ["f -4 -3 -2", "f -4/1/1 -2/1/1 -1/1/1 -3"]          # relative; a fourth corner is ignored
        else:
            base = 3 + 4 * q
            lines += ["f %d//7 %d//7 %d//7" % (base + 1, base + 2, base + 3), "f +%d %d %d" % (base + 1, base + 3, base + 4)]
        if bad_at is not None and q == bad_at:
            lines.append("f 1 2 x")
        if forward_ref_at is not None and q == forward_ref_at:
            lines.append("f 1 2 %d" % (3 + 4 * (q + 1) + 1))                # a vertex that is defined only later
        if out_of_range_at is not None and q == out_of_range_at:
            lines.append("f 1 2 %d" % (3 + 4 * n_quads + 1))               # one past the last vertex of the file
    with open(path, "w") as f:
        f.write("\n".join(lines) + "\n")
    return lines


def test_chunked_ingest_is_independent_of_the_thread_count(crt, orc, tmp_path, monkeypatch):
    (tmp_path / "big.mtl").write_text("newmtl m0\nKd .5 .5 .5\nnewmtl m1\nKd .1 .2 .3\nKe 5 5 5\nnewmtl m2\nKd .3 .2 .1\nNs 40\n")
    (tmp_path / "wrong.mtl").write_text("newmtl m0\nKd 1 0 0\n")
    for ref_forms in (False, True):
        obj = str(tmp_path / ("big%d.obj" % ref_forms))
        _big_obj(obj, reference_forms_only=ref_forms)
        assert os.path.getsize(obj) > (1 << 20)
        want = None
        for th in ("1", "2", "5", "8"):
            monkeypatch.setenv("CRT_INGEST_THREADS", th)
            S = crt.Scene().add_obj(obj, str(tmp_path))
            tr = S.tris()
            blob = b"".join(np.ascontiguousarray(tr[k]).tobytes() for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj")) + S.mats().tobytes()
            if want is None:
                want = blob
                assert S.counts()["n_tris"] == 24000 and len(S.lights()) == 1 and S.counts()["n_mats"] == 4
                if ref_forms:                   # the restatement of the reference's reader gives the same scene
                    O = orc.Scene().add_obj(obj, str(tmp_path))
                    ot = O.tris()
                    for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj"):
                        assert np.array_equal(tr[k].view(np.uint32), ot[k].view(np.uint32)), k
                    assert np.array_equal(S.mats(), O.mats()) and len(O.lights()) == 1
            assert blob == want, th


def test_chunked_ingest_reports_the_first_error_with_its_line(crt, tmp_path, monkeypatch):
    (tmp_path / "big.mtl").write_text("newmtl m0\nKd .5 .5 .5\n")
    (tmp_path / "wrong.mtl").write_text("newmtl m0\nKd 1 0 0\n")
    for kind in ("bad", "range"):
        obj = str(tmp_path / ("err_%s.obj" % kind))
        lines = _big_obj(obj, bad_at=9000 if kind == "bad" else 11000, out_of_range_at=7000 if kind == "range" else 10000)
        first = min(i for i, ln in enumerate(lines) if ln == "f 1 2 x" or (ln.startswith("f 1 2 ") and ln != "f 1 2 3")) + 1
        msgs = set()
        for th in ("1", "3", "8"):
            monkeypatch.setenv("CRT_INGEST_THREADS", th)
            with pytest.raises(crt.CrtError) as e:
                crt.Scene().add_obj(obj, str(tmp_path))
            msgs.add(str(e.value))
        assert len(msgs) == 1 and (":%d: malformed face" % first) in msgs.pop()


def test_forward_references_are_accepted_like_the_reference(crt, orc, tmp_path, monkeypatch):
    """OBJLoader.h:106 resolves positive indices after the whole file is read: a face may name a vertex that is defined
    later (in another chunk, too); negative indices count back from the vertices read so far."""
    (tmp_path / "big.mtl").write_text("newmtl m0\nKd .5 .5 .5\nnewmtl m1\nKd .1 .2 .3\nKe 5 5 5\nnewmtl m2\nKd .3 .2 .1\nNs 40\n")
    (tmp_path / "wrong.mtl").write_text("newmtl m0\nKd 1 0 0\n")
    obj = str(tmp_path / "fwd.obj")
    _big_obj(obj, forward_ref_at=7000, reference_forms_only=True)     # the forms the reference's stoull-based reader takes
    O = orc.Scene().add_obj(obj, str(tmp_path))
    for th in ("1", "5"):
        monkeypatch.setenv("CRT_INGEST_THREADS", th)
        S = crt.Scene().add_obj(obj, str(tmp_path))
        assert S.counts()["n_tris"] == O.n_tris == 2 * 12000 + 1
        assert np.array_equal(S.tris()["verts"].view(np.uint32), O.tris()["verts"].view(np.uint32))


def test_failed_calls_leave_the_scene_unchanged(crt, tmp_path):
    """A rejected OBJ or triangle batch rolls back triangles, material slots, objects and lights; ids are validated."""
    (tmp_path / "m.mtl").write_text("newmtl a\nKd 1 1 1\nnewmtl glow\nKe 3 3 3\n")
    (tmp_path / "ok.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nusemtl a\nf 1 2 3\nusemtl glow\nf 3 2 1\n")
    (tmp_path / "nan.obj").write_text("mtllib m.mtl\nv 0 0 0\nv 1 0 0\nv 0 nan 0\nusemtl glow\nf 1 2 3\nusemtl a\nf 3 2 1\n")
    S = crt.Scene().add_obj(str(tmp_path / "ok.obj"), str(tmp_path))
    before = (S.counts(), S.tris()["verts"].tobytes(), S.mats().tobytes(), [(f.tolist(), a) for f, a in S.lights()])
    with pytest.raises(crt.CrtError):
        S.add_obj(str(tmp_path / "nan.obj"), str(tmp_path))
    tri = [[0, 0, 0, 1, 0, 0, 0, 1, 0]]
    mat = [[.5, .5, .5, 0, 0, 0, 1]]
    for bad in (dict(verts=[[0, 0, 0, 1, 0, 0, 0, float("inf"), 0]], mat_id=[0], obj_id=[0]),
                dict(verts=tri, mat_id=[0], obj_id=[1]),                      # ids number this call's groups: < n_tris
                dict(verts=tri, mat_id=[0], obj_id=[0xFFFFFFFF]),
                dict(verts=tri, mat_id=[1], obj_id=[0])):
        with pytest.raises(crt.CrtError):
            S.add_triangles(bad["verts"], bad["mat_id"], bad["obj_id"], mat)
    with pytest.raises(ValueError):
        S.add_triangles(tri + tri, [0], [0, 0], mat)                          # array lengths are checked by the binding
    after = (S.counts(), S.tris()["verts"].tobytes(), S.mats().tobytes(), [(f.tolist(), a) for f, a in S.lights()])
    assert before == after
    S.add_triangles(tri, [0], [0], mat)
    assert S.counts()["n_tris"] == 3 and S.counts()["n_mats"] == 3


def test_hostile_inputs_are_errors_not_aborts(crt, tmp_path):
    """A PNG header that promises 2^48 pixels behind a few bytes of IDAT and a config nested 10^5 deep come back as
    CrtError (no exception crosses the C ABI, nothing of the promised size is allocated)."""
    def chunk(typ, data):
        return struct.pack(">I", len(data)) + typ + data + struct.pack(">I", zlib.crc32(typ + data) & 0xffffffff)
    png = b"\x89PNG\r\n\x1a\n" + chunk(b"IHDR", struct.pack(">IIBBBBB", 1 << 24, 1 << 24, 8, 2, 0, 0, 0)) + \
        chunk(b"IDAT", zlib.compress(b"\0" * 64)) + chunk(b"IEND", b"")
    (tmp_path / "t.mtl").write_text("newmtl t\nKd 1 1 1\nmap_Kd bomb.png\n")
    (tmp_path / "bomb.png").write_bytes(png)
    (tmp_path / "t.obj").write_text("mtllib t.mtl\nv 0 0 0\nv 1 0 0\nv 0 1 0\nvt 0 0\nvt 1 0\nvt 0 1\nusemtl t\nf 1/1 2/2 3/3\n")
    with pytest.raises(crt.CrtError):
        crt.Scene().add_obj(str(tmp_path / "t.obj"), str(tmp_path))
    (tmp_path / "deep.json").write_text("[" * 100000 + "]" * 100000)
    with pytest.raises(crt.CrtError) as e:
        crt.load_config(str(tmp_path / "deep.json"))
    assert "nesting too deep" in str(e.value)


def test_png_writer_bands_do_not_depend_on_the_thread_count(crt, tmp_path, monkeypatch):
    """Frames above ~1 MB are deflated in bands on all host threads; the file is a function of the image only."""
    rng = np.random.default_rng(6)
    for (h, w) in [(92, 3840), (400, 1500), (1, 1)]:            # 2 bands (91 rows each at this width), 3 bands, 1 band
        img = np.clip(rng.integers(-20, 20, (h, w, 3)) + (np.arange(w)[None, :, None] * 255 // max(w - 1, 1)), 0, 255).astype(np.uint8)
        blobs = set()
        for th in ("1", "3", "8"):
            monkeypatch.setenv("CRT_INGEST_THREADS", th)
            p = str(tmp_path / ("t%s.png" % th))
            crt.write_png(p, img)
            blobs.add(open(p, "rb").read())
        assert len(blobs) == 1
        data = blobs.pop()
        pos, idat = 8, b""
        while pos < len(data):
            n, typ = struct.unpack(">I4s", data[pos:pos + 8])
            if typ == b"IDAT": idat += data[pos + 8:pos + 8 + n]
            pos += 12 + n
        rows = np.frombuffer(zlib.decompress(idat), np.uint8).reshape(h, 1 + w * 3)
        assert (rows[:, 0] == 0).all() and np.array_equal(rows[:, 1:].reshape(h, w, 3), img)


def test_chunked_ingest_fuzz_against_the_single_thread_pass(crt, tmp_path, monkeypatch):
    """Random small OBJ texts cut into many tiny chunks (CRT_INGEST_MIN_CHUNK forces it): the scene, or the error message
    with its line number, must equal the one-chunk pass - relative indices across chunk borders, faces before the
    first usemtl, usemtl / mtllib anywhere, malformed and out-of-range faces, blank lines, no trailing newline."""
    rng = np.random.default_rng(2024)
    (tmp_path / "f.mtl").write_text("newmtl a\nKd .5 .5 .5\nnewmtl b\nKd .2 .2 .2\nKe 3 3 3\nnewmtl c\nNs 50\nKd .1 .1 .1\n")
    mats = ["a", "b", "c", "missing"]

    def scene_blob(S):
        tr = S.tris()
        return b"".join(np.ascontiguousarray(tr[k]).tobytes() for k in ("verts", "normal", "area", "area_of_obj", "mat", "obj")) + \
            S.mats().tobytes() + repr([(f.tolist(), float(a)) for f, a in S.lights()]).encode()

    n_err = n_ok = 0
    for case in range(120):
        lines, nv = [], 0
        if rng.random() < 0.8:
            lines.append("mtllib f.mtl")
        for _ in range(int(rng.integers(5, 60))):
            r = rng.random()
            if r < 0.45:
                lines.append("v %.3f %.3f %.3f" % tuple(rng.uniform(-9, 9, 3)))
                nv += 1
            elif r < 0.50:
                lines.append("vt 0.1 0.2")
            elif r < 0.60:
                lines.append("usemtl " + mats[int(rng.integers(0, 4))])
            elif r < 0.63:
                lines.append("" if rng.random() < 0.5 else "# c")
            elif r < 0.65:
                lines.append("mtllib f.mtl")
            elif nv >= 3 or rng.random() < 0.1:
                kind = rng.random()
                if kind < 0.06:
                    lines.append("f 1 2")                                   # too few corners
                elif kind < 0.10:
                    lines.append("f 1 x 3")                                 # does not parse
                elif kind < 0.16:
                    lines.append("f 1 2 %d" % (nv + int(rng.integers(1, 4))))     # not read yet
                elif kind < 0.55 and nv >= 3:
                    lines.append("f %d %d %d" % tuple(-int(x) for x in rng.integers(1, nv + 1, 3)))
                elif nv >= 3:
                    a, b, c = (int(x) for x in rng.integers(1, nv + 1, 3))
                    lines.append("f %d/1/1 %d//2 +%d 1" % (a, b, c))
        text = "\n".join(lines) + ("" if rng.random() < 0.3 else "\n")
        obj = tmp_path / ("z%d.obj" % case)
        obj.write_text(text)
        results = []
        for th, mc in (("1", "1000000"), ("7", "16"), ("3", "40"), ("64", "1")):
            monkeypatch.setenv("CRT_INGEST_THREADS", th)
            monkeypatch.setenv("CRT_INGEST_MIN_CHUNK", mc)
            try:
                results.append(("ok", scene_blob(crt.Scene().add_obj(str(obj), str(tmp_path)))))
            except crt.CrtError as e:
                results.append(("err", str(e)))
        assert all(r == results[0] for r in results), (case, text, [r[0] if r[0] == "ok" else r for r in results])
        n_err += results[0][0] == "err"
        n_ok += results[0][0] == "ok"
    assert n_ok >= 20 and n_err >= 20            # both outcomes are exercised
