"""ORACLE — TEST INFRASTRUCTURE ONLY.

numpy statement of the reference's map_Kd rule (include/Loader.h:55-59,78-105): Kd of a triangle = mean of the three
texels at its corners' uv (uv of a corner = vt[VERTEX index], Loader.h:81-83). Texel arithmetic as written there,
including the swapped (x, y) it receives from stbi_load(path, &height, &width, ...): u scales by (image height - 1),
v by (image width - 1), byte offset = (v * image height + u) * channels, and bytes offset+0..2 are read whatever the
channel count is. Pinned against the real reference code by tests/golden/map_kd.npz (tools/texture_fixture.py explains
why only a same-uv scene can be pinned: the reference's own mean is undefined behaviour).
"""
import numpy as np


def frac01(x):
    x = np.asarray(x, np.float32)
    f = (x - np.trunc(x)).astype(np.float32)               # modff: fraction with the sign of x
    y = (f + np.float32(1.0)).astype(np.float32)
    return (y - np.trunc(y)).astype(np.float32)


def kd_from_texture(pixels, uv3):
    """pixels: (H, W, C) uint8 as stb_image returns them; uv3: (T, 3, 2) float32. Returns (T, 3) float32 Kd."""
    H, W, C = pixels.shape
    flat = np.concatenate([pixels.reshape(-1), np.zeros(4, np.uint8)])
    width, height = H, W                                   # the reference's variable names
    uv3 = np.asarray(uv3, np.float32)
    u = (frac01(uv3[:, :, 0]) * np.float32(width - 1)).astype(np.int64)
    v = (frac01(uv3[:, :, 1]) * np.float32(height - 1)).astype(np.int64)
    off = (v * width + u) * C
    tex = np.stack([flat[off + c] for c in range(3)], axis=-1).astype(np.float32) / np.float32(255.0)     # (T, 3 corners, 3)
    s = (tex[:, 0] + tex[:, 1]).astype(np.float32)
    s = (s + tex[:, 2]).astype(np.float32)
    return (s / np.float32(3.0)).astype(np.float32)
