"""Randomised soak of both paths against the CPU oracle: random cameras, lights, anti-aliasing / soft-shadow /
depth-of-field settings, screen sizes, row bands and scenes (Cornell box, tessellations, random soups), every output
array compared bit for bit.  Runs for B2R_SOAK_SECONDS (default 8 s, so the normal suite stays short; the round's
long run is recorded in profiles/).  Writes gpurun_out/soak.json when that directory exists."""
import json
import os
import time

import numpy as np
import pytest

from util import ROOT, bits, random_soup, rot_y

pytestmark = pytest.mark.gpu


def _random_rt_case(pkg, rng):
    w, h = int(rng.integers(17, 200)), int(rng.integers(9, 150))
    kind = rng.integers(0, 4)
    if kind == 0:
        tris = pkg.cornell_box()
    elif kind == 1:
        tris = pkg.tessellate(pkg.cornell_box(), int(rng.integers(2, 5)))  # multi-chunk tables
    elif kind == 2:
        tris = random_soup(rng, int(rng.integers(1, 32)), spread=0.9, size=0.8)
    else:
        tris = np.concatenate([pkg.cornell_box(), random_soup(rng, int(rng.integers(1, 80)), spread=0.7, size=0.5)])
    fp = pkg.default_frame_params(0, w, h)
    yaw = rng.uniform(-0.6, 0.6)
    pos = [float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-0.5, 0.5)), float(rng.uniform(-2.6, -1.2))]
    fp.set_camera(pos, rot_y(yaw), float(rng.uniform(0.3, 1.2) * h))
    nl = int(rng.integers(1, 4))
    lights = np.concatenate([rng.uniform(-0.8, 0.8, (nl, 3)), rng.uniform(0.2, 1, (nl, 3)), rng.uniform(2, 20, (nl, 1))], 1)
    fp.set_lights(lights.astype(np.float32))
    aa = int(rng.choice([0, 0, 2, 3, 4]))
    fp.aaEnabled, fp.aaSamples = int(aa > 0), max(aa, 1)
    if rng.random() < 0.3:
        fp.softShadowsEnabled = 1
        fp.softShadowsSamples = int(rng.choice([2, 4, 16]))
        fp.set_random_positions(pkg.jitter_table(int(rng.integers(1, 1000)), lights[0, :3].astype(np.float32)))
    fp.dofEnabled = int(rng.random() < 0.3)
    return tris, fp, w, h


def _random_ras_case(pkg, rng):
    w, h = int(rng.integers(17, 260)), int(rng.integers(9, 200))
    kind = rng.integers(0, 3)
    if kind == 0:
        tris = pkg.cornell_box()
    elif kind == 1:
        tris = pkg.tessellate(pkg.cornell_box(), int(rng.integers(2, 14)))
    else:
        tris = random_soup(rng, int(rng.integers(1, 120)), spread=1.1, size=0.9)
        tris[:, [2, 5, 8]] += 1.0  # keep every vertex in front of the camera
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera([float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-0.3, 0.3)), float(rng.uniform(-3.4, -2.6))],
                  rot_y(rng.uniform(-0.25, 0.25), 1.01), float(rng.uniform(0.6, 1.3) * h))
    nl = int(rng.integers(1, 4))
    lights = np.concatenate([rng.uniform(-1, 1, (nl, 3)), rng.uniform(0.2, 1, (nl, 3)), rng.uniform(2, 20, (nl, 1))], 1)
    fp.set_lights(lights.astype(np.float32))
    fp.backfaceCulling, fp.frustumCulling = int(rng.random() < 0.7), int(rng.random() < 0.7)
    fp.dofEnabled = int(rng.random() < 0.3)
    return tris, fp, w, h


def test_soak_random_frames(pkg, oracle):
    seconds = float(os.environ.get("B2R_SOAK_SECONDS", "8"))
    rng = np.random.default_rng(int(os.environ.get("B2R_SOAK_SEED", "2026")))
    t_end = time.time() + seconds
    n_rt = n_ras = pixels = 0
    while time.time() < t_end:
        # ---- raytracer
        tris, fp, w, h = _random_rt_case(pkg, rng)
        ctx = pkg.Context(w, h)
        ctx.set_option(pkg.capi.OPT_RT_VARIANT, int(rng.choice([0, 0, 0, 1, 2, 3, 5])))
        ctx.set_triangles(tris)
        ctx.set_frame(fp)
        want = oracle.rt_draw(tris, fp, w, h)
        y0 = int(rng.integers(0, h))
        y1 = int(rng.integers(y0, h + 1))
        for a, b in ((0, h), (y0, y1)):
            got = ctx.rt_draw(a, b)
            for k in ("pixelColours", "focalDistances"):
                assert np.array_equal(bits(got[k][a:b]), bits(want[k][a:b])), ("rt", k, n_rt)
            assert np.array_equal(got["closest"][a:b].view(np.uint8), want["closest"][a:b].view(np.uint8)), ("rt closest", n_rt)
        surf = oracle.resolve_surface(want["pixelColours"], want["focalDistances"], bool(fp.dofEnabled), 8)
        assert np.array_equal(ctx.rt_frame(), surf), ("rt frame", n_rt)
        ctx.close()
        n_rt += 1
        pixels += w * h
        # ---- rasteriser
        tris, fp, w, h = _random_ras_case(pkg, rng)
        ctx = pkg.Context(w, h)
        ctx.set_option(pkg.capi.OPT_RAS_VARIANT, int(rng.integers(0, 4)))  # both pipelines, with and without fixed slots
        ctx.set_triangles(tris)
        ctx.set_frame(fp)
        culled = ctx.ras_cull()
        assert np.array_equal(culled, oracle.ras_cull(tris, fp, w, h)), ("cull", n_ras)
        want = oracle.ras_draw(tris, culled, fp, w, h)
        y0 = int(rng.integers(0, h))
        y1 = int(rng.integers(y0, h + 1))
        for a, b in ((0, h), (y0, y1)):
            got = ctx.ras_draw(a, b)
            for k in ("depthBuffer", "pixelColours", "focalDistances", "winner"):
                assert np.array_equal(bits(got[k][a:b]), bits(want[k][a:b])), ("ras", k, n_ras)
        surf = oracle.resolve_surface(want["pixelColours"], want["focalDistances"], bool(fp.dofEnabled), 8)
        assert np.array_equal(ctx.ras_frame(), surf), ("ras frame", n_ras)
        ctx.close()
        n_ras += 1
        pixels += w * h
    summary = {"seconds": seconds, "raytracer_cases": n_rt, "rasteriser_cases": n_ras, "pixels": pixels,
               "mismatches": 0, "compared": "every output array, bit for bit, full frame + one random row band + the "
                                            "frame call's surface, CUDA (C ABI) vs oracle/liboracle.so"}
    out = os.path.join(ROOT, "gpurun_out")
    if os.path.isdir(out):
        with open(os.path.join(out, "soak.json"), "w") as f:
            json.dump(summary, f)
    assert n_rt > 0 and n_ras > 0
