"""Every BASELINE configuration at the size bench.py quotes it (3840x2160), full frame, against digests produced by
the REFERENCE ITSELF (oracle/_ref/libref_*_3840x2160.so, tests/golden/make_golden.py kat4k): config 3 (AA 4x4), config
3b (16 jittered light samples), three frames of the config-5 orbit, config 4 (1,004,670 triangles) -- the latter on both
rasteriser pipelines.  Digests only: nothing of the CPU side runs on the GPU box."""
import hashlib
import json
import os

import numpy as np
import pytest

from util import GOLDEN

pytestmark = pytest.mark.gpu
W, H = 3840, 2160


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def kat():
    return json.load(open(os.path.join(GOLDEN, "kat_3840x2160.json")))


def _check_rt(ctx, want):
    got = ctx.rt_draw()
    for name, key in (("pixelColours", "pixelColours"), ("focalDistances", "focalDistances"), ("closest", "closest")):
        assert sha(got[key]) == want["sha256"][name], name
    assert sha(ctx.rt_frame()) == want["sha256"]["surface"], "surface"
    return got


def test_config3_aa4x4_full_frame(pkg, kat):
    ctx = pkg.Context(W, H)
    ctx.set_triangles(pkg.cornell_box())
    fp = pkg.default_frame_params(0, W, H)
    fp.aaEnabled, fp.aaSamples = 1, 4
    ctx.set_frame(fp)
    got = _check_rt(ctx, kat["config3_aa4x4"])
    assert int((got["closest"]["triangleIndex"] >= 0).sum()) == kat["config3_aa4x4"]["hit_pixels"]
    # the split forms of the same frame: gather of 8 parts on one device, and a group of every GPU of the box
    import torch
    dev = torch.device("cuda:0")
    root = torch.zeros((H * W + 64,), dtype=torch.int32, device=dev)
    for part in range(8):
        ctx.rt_frame_gather_device_async(part, 8, root.data_ptr(), root.data_ptr() + H * W * 4)
    ctx.synchronize()
    assert int(root[H * W].item()) == 8
    assert sha(root[:H * W].cpu().numpy()) == kat["config3_aa4x4"]["sha256"]["surface"], "gathered surface"
    ctx.close()
    g = pkg.Group(W, H, list(range(torch.cuda.device_count())))
    g.set_triangles(pkg.cornell_box())
    g.set_frame(fp)
    assert sha(g.rt_frame()) == kat["config3_aa4x4"]["sha256"]["surface"], "group surface"
    g.close()


def test_config3b_soft_shadows_full_frame(pkg, kat):
    ctx = pkg.Context(W, H)
    ctx.set_triangles(pkg.cornell_box())
    fp = pkg.default_frame_params(0, W, H)
    fp.softShadowsEnabled, fp.softShadowsSamples = 1, 16
    table = pkg.jitter_table(1, [0, -0.5, -0.7])
    assert sha(np.asarray(table, np.float32).reshape(256, 3)) == kat["config3b_soft16"]["jitter_table_sha256"]
    fp.set_random_positions(table)
    ctx.set_frame(fp)
    _check_rt(ctx, kat["config3b_soft16"])
    ctx.close()


def test_config5_orbit_frames_full_size(pkg, kat):
    ctx = pkg.Context(W, H)
    ctx.set_triangles(pkg.cornell_box())
    for f, want in kat["config5_orbit"].items():
        fp = pkg.default_frame_params(0, W, H)
        pos, rot = pkg.orbit_camera(int(f), 360)
        fp.set_camera(pos, rot, H / 2)
        ctx.set_frame(fp)
        _check_rt(ctx, want)
    ctx.close()


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_config4_rasteriser_full_frame(pkg, kat, variant):
    want = kat["config4_ras_1m"]
    tris = pkg.tessellate(pkg.cornell_box(), 183)
    assert len(tris) == want["triangles"]
    ctx = pkg.Context(W, H)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, variant)
    ctx.enable_stats(True)
    ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(1, W, H))
    culled = ctx.ras_cull()
    assert int(culled.sum()) == want["culled"] and sha(culled.astype(np.uint8)) == want["culled_sha256"]
    got = ctx.ras_draw()
    assert ctx.stats()["ras_depth_tests"] == want["depth_tests"]
    assert int((got["winner"] >= 0).sum()) == want["covered"]
    for name in ("depthBuffer", "pixelColours", "focalDistances", "winner"):
        assert sha(got[name]) == want["sha256"][name], name
    assert sha(ctx.ras_frame()) == want["sha256"]["surface"], "surface"
    ctx.close()
