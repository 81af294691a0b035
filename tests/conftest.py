import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

REFERENCE = "/root/reference"
HAS_REFERENCE = os.path.isdir(os.path.join(REFERENCE, "scenes"))
GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")
    config.addinivalue_line("markers", "reference: needs the read-only reference checkout at /root/reference")


def pytest_collection_modifyitems(config, items):
    skip_ref = pytest.mark.skip(reason="/root/reference not present on this machine")
    for item in items:
        if "reference" in item.keywords and not HAS_REFERENCE:
            item.add_marker(skip_ref)


@pytest.fixture(scope="session")
def crt():
    import cudaraytracing_b200 as m
    m.build_native()
    m.load_library()
    return m


@pytest.fixture(scope="session")
def orc():
    from oracle import orc as m
    m.lib()
    return m


@pytest.fixture(scope="session")
def scene_files(tmp_path_factory):
    """name -> dict(cfg_path, obj, dir): the fixtures written back as OBJ/MTL/config.json."""
    from tools import scene_fixture as sf
    out = {}
    for name in ("cornell-box", "veach-mis"):
        d = str(tmp_path_factory.mktemp(name.replace("-", "_")))
        cfg_path = sf.unpack(sf.fixture(name), d)
        out[name] = dict(cfg_path=cfg_path, obj=os.path.join(d, name + ".obj"), dir=d)
    return out


def soup(rng, n, extent=10.0, size=1.0):
    """n random triangles; returns verts (n,9) float32."""
    c = rng.uniform(-extent, extent, (n, 1, 3))
    v = c + rng.uniform(-size, size, (n, 3, 3))
    return v.reshape(n, 9).astype(np.float32)


def box_scene(rng, n_soup=500):
    """A closed diffuse box with a ceiling light, a specular floor patch and a soup of small triangles.
    Returns verts, mat_id, obj_id, mats (kd[3], ke[3], ns)."""
    def quad(a, b, c, d):
        return [a + b + c, a + c + d]
    L = 10.0
    P = lambda x, y, z: [float(x), float(y), float(z)]
    tris, mat, obj = [], [], []
    walls = [
        (quad(P(0, 0, 0), P(L, 0, 0), P(L, 0, L), P(0, 0, L)), 0),      # floor
        (quad(P(0, L, 0), P(0, L, L), P(L, L, L), P(L, L, 0)), 0),      # ceiling
        (quad(P(0, 0, L), P(L, 0, L), P(L, L, L), P(0, L, L)), 0),      # back
        (quad(P(0, 0, 0), P(0, 0, L), P(0, L, L), P(0, L, 0)), 1),      # left
        (quad(P(L, 0, 0), P(L, L, 0), P(L, L, L), P(L, 0, L)), 2),      # right
    ]
    o = 0
    for q, m in walls:
        for t in q:
            tris.append(t); mat.append(m); obj.append(o)
        o += 1
    for t in quad(P(4, L - 0.01, 4), P(6, L - 0.01, 4), P(6, L - 0.01, 6), P(4, L - 0.01, 6)):   # light, facing down
        tris.append(t); mat.append(3); obj.append(o)
    o += 1
    for t in quad(P(2, 0.01, 2), P(2, 0.01, 8), P(8, 0.01, 8), P(8, 0.01, 2)):                   # glossy plate, facing up
        tris.append(t); mat.append(4); obj.append(o)
    o += 1
    s = soup(rng, n_soup, extent=3.0, size=0.4) + np.tile(np.array([5, 4, 6], np.float32), 3)
    for t in s:
        tris.append(list(map(float, t))); mat.append(5); obj.append(o)
    verts = np.array(tris, np.float32)
    mats = np.array([[0.7, 0.7, 0.7, 0, 0, 0, 1], [0.8, 0.1, 0.1, 0, 0, 0, 1], [0.1, 0.8, 0.1, 0, 0, 0, 1],
                     [0, 0, 0, 20, 18, 15, 1], [0.2, 0.3, 0.4, 0, 0, 0, 200], [0.6, 0.6, 0.3, 0, 0, 0, 1]], np.float32)
    return verts, np.array(mat, np.int32), np.array(obj, np.int32), mats


BOX_CAMERA = dict(eye=[5.0, 5.0, -12.0], lookat=[5.0, 4.5, 0.0], up=[0.0, 1.0, 0.0], fovy=math.radians(40.0))


def random_rays(rng, lo, hi, n, tmax_any=False):
    r = np.zeros((n, 8), np.float32)
    r[:, 0:3] = rng.uniform(lo, hi, (n, 3))
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r[:, 4:7] = d
    r[:, 3] = rng.uniform(0, np.linalg.norm(np.asarray(hi) - np.asarray(lo)), n) if tmax_any else np.finfo(np.float32).max
    return r
