"""Compact scene fixtures.

The reference's scenes (scenes/cornell-box, scenes/veach-mis: OBJ + MTL + config.json) are the
inputs of every BASELINE config, but /root/reference does not exist on the GPU box. `pack` (run
here, by tools/make_golden.py) stores what the loader consumes — vertex positions, the first three
vertex indices of every face, the usemtl groups, the MTL text and the config — as an lzma-compressed
npz; `unpack` writes it back as an equivalent OBJ/MTL/config.json triple in a scratch directory so
that tests and bench.py go through the product's real OBJ/MTL/JSON ingest. Equivalence (identical
triangles, materials and lights after parsing) is asserted by tests/test_ingest.py.
"""
import io
import json
import lzma
import os

import numpy as np


def parse_obj_minimal(obj_path):
    verts, faces, groups = [], [], []
    mtllib = None
    with open(obj_path) as f:
        for line in f:
            if line.startswith("v "):
                p = line.split()
                verts.append((p[1], p[2], p[3]))
            elif line.startswith("f "):
                p = line.split()[1:4]
                faces.append([int(x.split("/")[0]) for x in p])
            elif line.startswith("usemtl"):
                groups.append((len(faces), line.split()[1]))
            elif line.startswith("mtllib"):
                mtllib = line.split()[1]
    v = np.array(verts, dtype=np.float32)
    f = np.array(faces, dtype=np.int64) - 1
    return v, f, groups, mtllib


def pack(obj_path, mtl_dir, config_path, out_path):
    v, f, groups, mtllib = parse_obj_minimal(obj_path)
    uniq, inv = np.unique(v, axis=0, return_inverse=True)
    f = inv.reshape(-1)[f].astype(np.uint32)
    with open(os.path.join(mtl_dir, mtllib)) as fh:
        mtl_text = fh.read()
    with open(config_path) as fh:
        config = json.load(fh)
    meta = dict(groups=[[int(s), n] for s, n in groups], mtllib=mtllib, mtl_text=mtl_text, config=config,
                name=os.path.splitext(os.path.basename(obj_path))[0])
    buf = io.BytesIO()
    np.savez(buf, verts=uniq.astype(np.float32), faces=f, meta=np.frombuffer(json.dumps(meta).encode(), dtype=np.uint8))
    with open(out_path, "wb") as fh:
        fh.write(lzma.compress(buf.getvalue(), preset=9))
    return out_path


def load(fixture_path):
    with open(fixture_path, "rb") as fh:
        data = np.load(io.BytesIO(lzma.decompress(fh.read())))
    meta = json.loads(bytes(data["meta"]).decode())
    return data["verts"], data["faces"], meta


def unpack(fixture_path, out_dir):
    """Writes <name>.obj, the MTL and config.json into out_dir; returns the config path."""
    verts, faces, meta = load(fixture_path)
    os.makedirs(out_dir, exist_ok=True)
    name = meta["name"]
    lines = ["mtllib %s\n" % meta["mtllib"]]
    # every vertex gets a vn and a vt so that "a/a/a" corners stay valid for the reference's parser
    # (include/OBJLoader.h:98-118 reads all three indices; Loader.h:70-72 indexes normals by vertex)
    for x, y, z in verts:
        lines.append("v %s %s %s\nvn 0 1 0\nvt 0 0\n" % ("%.9g" % x, "%.9g" % y, "%.9g" % z))
    bounds = [s for s, _ in meta["groups"]] + [len(faces)]
    for gi, (start, mat) in enumerate(meta["groups"]):
        lines.append("g\nusemtl %s\n" % mat)
        for a, b, c in faces[start:bounds[gi + 1]] + 1:
            lines.append("f %d/%d/%d %d/%d/%d %d/%d/%d\n" % (a, a, a, b, b, b, c, c, c))
    with open(os.path.join(out_dir, name + ".obj"), "w") as fh:
        fh.writelines(lines)
    with open(os.path.join(out_dir, meta["mtllib"]), "w") as fh:
        fh.write(meta["mtl_text"])
    cfg = dict(meta["config"])
    cfg["OBJ_paths"] = [{"OBJ_path": name + ".obj", "MTL_dir": "."}]
    cfg_path = os.path.join(out_dir, "config.json")
    with open(cfg_path, "w") as fh:
        json.dump(cfg, fh, indent=1)
    return cfg_path


FIXTURE_DIR = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden", "scenes")


def fixture(name):
    return os.path.join(FIXTURE_DIR, name + ".npz.xz")
