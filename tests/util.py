"""Shared helpers for the parity tests."""
import os

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLDEN = os.path.join(ROOT, "tests", "golden")
HAVE_REFERENCE_SOURCES = os.path.isdir("/root/reference")

DEFAULT_LIGHT = [[0, -0.5, -0.7, 1, 1, 1, 14]]


def fnv1a32(data: bytes) -> int:
    h = 0x811C9DC5
    for b in data:
        h = ((h ^ b) * 0x01000193) & 0xFFFFFFFF
    return h


def bits(a):
    return np.ascontiguousarray(a).view(np.uint32)


def random_soup(rng, n, spread=1.0, size=0.6):
    """n random triangles (raytracer layout, normals like the Triangle ctor) around the origin."""
    v0 = rng.uniform(-spread, spread, (n, 3)).astype(np.float32)
    v1 = (v0 + rng.uniform(-size, size, (n, 3))).astype(np.float32)
    v2 = (v0 + rng.uniform(-size, size, (n, 3))).astype(np.float32)
    col = rng.uniform(0.1, 0.9, (n, 3)).astype(np.float32)
    e1, e2 = v1 - v0, v2 - v0
    nrm = np.cross(e2, e1).astype(np.float32)
    ln = np.sqrt((nrm * nrm).sum(1, keepdims=True)).astype(np.float32)
    ln[ln == 0] = 1
    nrm = (nrm / ln).astype(np.float32)
    return np.concatenate([v0, v1, v2, nrm, col], axis=1).astype(np.float32)


def rot_y(yaw, r11=1.0):
    c, s = np.float32(np.cos(np.float32(yaw))), np.float32(np.sin(np.float32(yaw)))
    return np.array([c, 0, s, 0, r11, 0, -s, 0, c], np.float32)


def rt_compare(got, want, tol=1e-4):
    """North-star criteria for the raytracer; returns a dict of counts."""
    gi, wi = got["closest"]["triangleIndex"], want["closest"]["triangleIndex"]
    n = gi.size
    idx_mismatch = int((gi != wi).sum())
    col_err = float(np.abs(got["pixelColours"] - want["pixelColours"]).max())
    return dict(pixels=n, idx_mismatch=idx_mismatch, idx_match_frac=1.0 - idx_mismatch / n, colour_max_abs=col_err,
                colours_bit_equal=bool(np.array_equal(bits(got["pixelColours"]), bits(want["pixelColours"]))),
                closest_bit_equal=bool(np.array_equal(got["closest"].view(np.uint8), want["closest"].view(np.uint8))),
                focal_bit_equal=bool(np.array_equal(bits(got["focalDistances"]), bits(want["focalDistances"]))))


def golden_files(prefix):
    import glob
    return sorted(glob.glob(os.path.join(GOLDEN, prefix + "*.npz")))


def rt_params_from_golden(pkg, z, w, h):
    fp = pkg.default_frame_params(0, w, h)
    fp.set_camera(z["pos"], z["rot"], float(z["focal"]))
    aa, soft = int(z["aa"]), int(z["soft"])
    fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled = int(aa > 0), max(aa, 1), soft
    if "soft_samples" in z:
        fp.softShadowsSamples = int(z["soft_samples"])
    if "lights" in z:
        fp.set_lights(z["lights"])
    fp.set_random_positions(z["table"])
    return fp


def ras_params_from_golden(pkg, z, w, h):
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera(z["pos"], z["rot"], float(z["focal"]))
    return fp


def size_from_name(path):
    import re
    m = re.search(r"_(\d+)x(\d+)_", os.path.basename(path))
    return int(m.group(1)), int(m.group(2))
