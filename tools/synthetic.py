"""Procedural scenes and ray sets of BASELINE configs C4 / C5; generated from a seed, never stored.

C4: a box [0,1000]^3 open towards the camera (5 walls, 10 triangles, diffuse 0.7) + one 200x200 quad light under the
ceiling (2 triangles, Ke 40) + a heightfield over the floor with n x n vertices -> 2 (n-1)^2 triangles; n = 2237
gives 9,999,392 + 12 = 9,999,404 triangles. height = 60 * (three sin/cos octaves) + 5 * Philox(seed 0x5EED, vertex id).
Camera eye (500,500,-1400) -> lookat (500,300,500), fov_y 40 deg, 1920x1080, spp 64, lsn 1, P_RR 0.6.
C5: rays with origin uniform in the scene box and direction uniform on the sphere, Philox key 0xC5, counter = ray
index (generated on the device by crt_random_rays_device; `random_rays` is the numpy statement for small n).
"""
import math

import numpy as np

C4_CAMERA = dict(eye=[500.0, 500.0, -1400.0], lookat=[500.0, 300.0, 500.0], up=[0.0, 1.0, 0.0], fov_y=40.0,
                 width=1920, height=1080, spp=64, light_sample_n=1, P_RR=0.6, bvh_thresh_n=2)
C4_FULL_N = 2237


def philox4x32_10(c0, c1, c2, c3, k0, k1):
    """Vectorised Philox4x32-10 on uint32 arrays (Salmon et al. 2011)."""
    c = [np.asarray(x, np.uint64) for x in (c0, c1, c2, c3)]
    k0, k1 = np.uint64(k0), np.uint64(k1)
    M0, M1, mask = np.uint64(0xD2511F53), np.uint64(0xCD9E8D57), np.uint64(0xFFFFFFFF)
    for _ in range(10):
        p0, p1 = M0 * c[0], M1 * c[2]
        c = [((p1 >> np.uint64(32)) ^ c[1] ^ k0) & mask, p1 & mask, ((p0 >> np.uint64(32)) ^ c[3] ^ k1) & mask, p0 & mask]
        k0 = (k0 + np.uint64(0x9E3779B9)) & mask
        k1 = (k1 + np.uint64(0xBB67AE85)) & mask
    return [x.astype(np.uint32) for x in c]


def u01(x):
    return ((x >> np.uint32(8)) + np.uint32(1)).astype(np.float32) * np.float32(2.0 ** -24)


def c4_scene(n=C4_FULL_N, seed=0x5EED):
    """Returns verts (T,9) float32, mat_id, obj_id (uint32), mats (3,7) = kd, ke, ns."""
    L = np.float32(1000.0)
    xs = np.linspace(0.0, 1000.0, n, dtype=np.float32)
    X, Z = np.meshgrid(xs, xs, indexing="xy")                      # [row z, col x]
    vid = np.arange(n * n, dtype=np.uint32)
    zero = np.zeros_like(vid)
    noise = u01(philox4x32_10(vid, zero, zero, zero, seed, 0)[0]).reshape(n, n)
    f = np.float32(2 * math.pi / 1000.0)
    Y = (np.float32(60.0) * (np.sin(f * X * 2) * np.cos(f * Z * 3) * np.float32(0.5)
                             + np.sin(f * X * 5 + 1) * np.cos(f * Z * 7) * np.float32(0.3)
                             + np.sin(f * X * 11) * np.sin(f * Z * 13 + 2) * np.float32(0.2)) + np.float32(5.0) * noise + np.float32(80.0)).astype(np.float32)
    P = np.stack([X, Y, Z], axis=-1).astype(np.float32)             # (n, n, 3)
    a, b, c, d = P[:-1, :-1], P[:-1, 1:], P[1:, 1:], P[1:, :-1]     # quad corners; normals point up (+y)
    t1 = np.concatenate([a, d, c], axis=-1).reshape(-1, 9)
    t2 = np.concatenate([a, c, b], axis=-1).reshape(-1, 9)
    field = np.empty((t1.shape[0] * 2, 9), np.float32)
    field[0::2], field[1::2] = t1, t2

    def quad(p0, p1, p2, p3):
        return [p0 + p1 + p2, p0 + p2 + p3]
    l = float(L)
    walls = (quad([0, 0, 0], [l, 0, 0], [l, 0, l], [0, 0, l]) + quad([0, l, 0], [0, l, l], [l, l, l], [l, l, 0]) +
             quad([0, 0, l], [l, 0, l], [l, l, l], [0, l, l]) + quad([0, 0, 0], [0, 0, l], [0, l, l], [0, l, 0]) +
             quad([l, 0, 0], [l, l, 0], [l, l, l], [l, 0, l]))
    light = quad([400, l - 1, 400], [600, l - 1, 400], [600, l - 1, 600], [400, l - 1, 600])
    fixed = np.array(walls + light, np.float32).reshape(-1, 3, 3)
    nrm = np.cross(fixed[:, 1] - fixed[:, 0], fixed[:, 2] - fixed[:, 0])
    outward = np.einsum("ij,ij->i", nrm, np.array([500.0, 500.0, 500.0], np.float32) - fixed[:, 0]) < 0
    fixed[outward] = fixed[outward][:, [0, 2, 1]]                  # all walls and the light face the inside of the box
    verts = np.concatenate([fixed.reshape(-1, 9), field], axis=0)
    mat = np.concatenate([np.zeros(10, np.uint32), np.full(2, 1, np.uint32), np.full(len(field), 2, np.uint32)])
    obj = np.concatenate([np.repeat(np.arange(5, dtype=np.uint32), 2), np.full(2, 5, np.uint32), np.full(len(field), 6, np.uint32)])
    mats = np.array([[0.7, 0.7, 0.7, 0, 0, 0, 1], [0, 0, 0, 40, 40, 40, 1], [0.55, 0.6, 0.45, 0, 0, 0, 1]], np.float32)
    return verts, mat, obj, mats


def _sincos_2pi_vec(u):
    """Vectorised sincos_2pi (crt_device.cuh) for large n: fmaf is emulated in float64 (exact product, the sum rounded
    twice), so a direction may differ from the device's in the last bit on rare inputs. Used for the timing baselines
    of the reference arm only; the parity tests use the exact scalar statement (n <= 200000)."""
    f32 = np.float32
    def fma(a, b, c):
        return (a.astype(np.float64) * np.asarray(b, np.float64) + np.asarray(c, np.float64)).astype(f32)
    u = u.astype(f32)
    q = fma(u, f32(4.0), f32(0.5)).astype(np.int32)
    x = ((u - q.astype(f32) * f32(0.25)) * f32(6.2831855)).astype(f32)
    x2 = (x * x).astype(f32)
    ps = fma(x2, f32(2.7557319e-06), f32(-1.9841270e-04))
    ps = fma(ps, x2, f32(8.3333333e-03))
    ps = fma(ps, x2, f32(-1.6666667e-01))
    ss = fma((x * x2).astype(f32), ps, x)
    pc = fma(x2, f32(-2.7557319e-07), f32(2.4801587e-05))
    pc = fma(pc, x2, f32(-1.3888889e-03))
    pc = fma(pc, x2, f32(4.1666667e-02))
    pc = fma(pc, x2, f32(-0.5))
    cc = fma(pc, x2, f32(1.0))
    k = q & 3
    s = np.where(k == 0, ss, np.where(k == 1, cc, np.where(k == 2, -ss, -cc)))
    c = np.where(k == 0, cc, np.where(k == 1, -ss, np.where(k == 2, -cc, ss)))
    return np.stack([s, c], axis=1).astype(f32)


def random_rays(lo, hi, n, key=0xC5, any_hit=False, start=0):
    """numpy statement of crt_random_rays_device: n x 8 float32 {o, tmax, d, 0}."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    lo32, hi32 = (idx & np.uint64(0xFFFFFFFF)).astype(np.uint32), (idx >> np.uint64(32)).astype(np.uint32)
    zero = np.zeros(n, np.uint32)
    a = philox4x32_10(lo32, hi32, zero, zero, key, 0)
    b = philox4x32_10(lo32, hi32, zero + np.uint32(1), zero, key, 0)
    lo, hi = np.asarray(lo, np.float32), np.asarray(hi, np.float32)
    r = np.zeros((n, 8), np.float32)
    for k in range(3):
        r[:, k] = lo[k] + (hi[k] - lo[k]) * u01(a[k])
    z = np.float32(1.0) - np.float32(2.0) * u01(a[3])
    rad = np.sqrt(np.maximum(np.float32(0.0), np.float32(1.0) - z * z)).astype(np.float32)
    if n <= 200000:
        from oracle import orc
        sc = np.array([orc.sincos_2pi(float(u)) for u in u01(b[0])], np.float32)
    else:
        sc = _sincos_2pi_vec(u01(b[0]))
    r[:, 4], r[:, 5], r[:, 6] = rad * sc[:, 1], rad * sc[:, 0], z
    diag = np.float32(np.sqrt(np.sum((hi - lo).astype(np.float64) ** 2)))
    r[:, 3] = u01(b[1]) * diag if any_hit else np.finfo(np.float32).max
    return r
