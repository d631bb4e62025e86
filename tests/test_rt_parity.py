"""Raytracer parity: CUDA path (through the C ABI) vs the CPU oracle.

Criteria (BASELINE.json north_star): hit-triangle index bit-exact on >= 99.99 % of pixels, pixel
colour within 1e-4 absolute.  This implementation is in fact bit-exact on every output array, and
the tests assert that, reporting the north-star numbers alongside.
"""
import numpy as np
import pytest

from util import DEFAULT_LIGHT, bits, random_soup, rot_y, rt_compare

pytestmark = pytest.mark.gpu

COLOUR_TOL = 1e-4       # north star: absolute
INDEX_MATCH_MIN = 0.9999  # north star: fraction of pixels


def check(got, want, exact=True):
    r = rt_compare(got, want, COLOUR_TOL)
    assert r["idx_match_frac"] >= INDEX_MATCH_MIN, r
    assert r["colour_max_abs"] <= COLOUR_TOL, r
    if exact:
        assert r["idx_mismatch"] == 0, r
        assert r["closest_bit_equal"] and r["colours_bit_equal"] and r["focal_bit_equal"], r
    return r


def draw_both(pkg, oracle, tris, fp, w, h, filt=1, ctx=None):
    own = ctx is None
    ctx = ctx or pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RT_FILTER, filt)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    got = ctx.rt_draw()
    if own:
        ctx.close()
    return got, oracle.rt_draw(tris, fp, w, h)


@pytest.mark.parametrize("filt", [1, 0])
def test_config1_cornell_500(pkg, oracle, filt):
    """BASELINE config 1: Cornell box 500x500, primary + shadow ray."""
    w = h = 500
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    got, want = draw_both(pkg, oracle, tris, fp, w, h, filt)
    check(got, want)
    # known answers of the reference itself
    idx = got["closest"]["triangleIndex"]
    assert (idx < 0).sum() == 0
    hist = dict(zip(*[a.tolist() for a in np.unique(idx, return_counts=True)]))
    assert hist == {0: 35312, 1: 5084, 2: 41918, 3: 13862, 4: 40779, 5: 13662, 6: 41911, 7: 14016, 8: 4135,
                    9: 13153, 10: 6042, 11: 6944, 18: 1322, 19: 1026, 20: 5730, 21: 4837, 26: 154, 27: 113}


@pytest.mark.parametrize("w,h", [(96, 64), (64, 96), (160, 120), (33, 17), (1, 1)])
@pytest.mark.parametrize("aa,soft", [(0, 0), (3, 0), (0, 1), (2, 1), (4, 0)])
def test_modes_small(pkg, oracle, w, h, aa, soft):
    """AA sub-sampling quirks (x1 only advances on a hit, state carried between sub-samples) and soft shadows."""
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = int(aa > 0), max(aa, 1)
    fp.softShadowsEnabled = soft
    fp.set_random_positions(pkg.jitter_table(1, [0, -0.5, -0.7]))
    for filt in (1, 0):
        got, want = draw_both(pkg, oracle, tris, fp, w, h, filt)
        check(got, want)


@pytest.mark.parametrize("seed", [1, 2, 3])
def test_random_scenes_multi_light(pkg, oracle, seed):
    """Random triangle soups, rotated camera, several lights (result2 += result accumulation, raytracer.cpp:322)."""
    rng = np.random.default_rng(seed)
    w, h = 128, 72
    tris = random_soup(rng, 24 + 8 * seed)
    fp = pkg.default_frame_params(0, w, h)
    fp.set_camera(rng.uniform(-0.3, 0.3, 3).astype(np.float32) + np.array([0, 0, -2.5], np.float32),
                  rot_y(rng.uniform(-0.4, 0.4)), h / 2)
    lights = np.concatenate([rng.uniform(-1, 1, (3, 3)), rng.uniform(0.2, 1, (3, 3)), rng.uniform(2, 20, (3, 1))], 1)
    fp.set_lights(lights.astype(np.float32))
    fp.softShadowsSamples = 4
    fp.set_random_positions(rng.uniform(-1, 1, (256, 3)).astype(np.float32))
    for aa, soft in ((0, 0), (2, 0), (0, 1)):
        fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled = int(aa > 0), max(aa, 1), soft
        for filt in (1, 0):
            got, want = draw_both(pkg, oracle, tris, fp, w, h, filt)
            check(got, want)


def test_degenerate_inputs(pkg, oracle):
    """No lights; a scene with degenerate (zero-area) and camera-coplanar triangles; empty scene."""
    w, h = 64, 48
    rng = np.random.default_rng(7)
    tris = random_soup(rng, 12)
    tris[3, 3:9] = tris[3, 0:3].tolist() * 2            # zero-area triangle
    tris[5, 2] = tris[5, 5] = tris[5, 8] = -2.0           # plane through the camera z
    tris[6, [1, 4, 7]] = 0.0                              # plane y = 0 contains the camera
    fp = pkg.default_frame_params(0, w, h)
    got, want = draw_both(pkg, oracle, tris, fp, w, h)
    check(got, want)
    fp.numLights = 0
    got, want = draw_both(pkg, oracle, tris, fp, w, h)
    check(got, want)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(np.zeros((0, 15), np.float32))
    ctx.set_frame(pkg.default_frame_params(0, w, h))
    got = ctx.rt_draw()
    assert (got["closest"]["triangleIndex"] == -1).all() and not got["pixelColours"].any()
    ctx.close()


def test_row_bands_equal_full_frame(pkg, oracle):
    """Multi-GPU split: bands rendered separately land in the same offsets and equal the full frame."""
    w, h = 160, 120
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 2
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    full = ctx.rt_draw()
    merged = {k: np.zeros_like(v) for k, v in full.items()}
    for y0, y1 in ((0, 37), (37, 37), (37, 100), (100, 120)):
        part = ctx.rt_draw(y0, y1)
        for k in merged:
            merged[k][y0:y1] = part[k][y0:y1]
    for k in full:
        assert np.array_equal(full[k].view(np.uint8), merged[k].view(np.uint8)), k
    ctx.close()


@pytest.mark.parametrize("frame", [0, 57, 180, 301])
def test_orbit_frames(pkg, oracle, frame):
    """BASELINE config 5: camera orbit; frames partition across GPUs, each is an independent Draw()."""
    w = h = 200
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    pos, rot = pkg.orbit_camera(frame, 360)
    fp.set_camera(pos, rot, h / 2)
    got, want = draw_both(pkg, oracle, tris, fp, w, h)
    r = check(got, want)
    assert r["pixels"] == w * h


def test_config3_4k_single_sample_and_aa(pkg, oracle):
    """BASELINE config 3 at full size: 3840x2160; 1 spp bit-exact vs the oracle, and AA 4x4 (16 spp) on a row band."""
    w, h = 3840, 2160
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    got = ctx.rt_draw()
    want = oracle.rt_draw(tris, fp, w, h)
    r = check(got, want)
    miss = int((got["closest"]["triangleIndex"] < 0).sum())
    assert miss == 3626640, miss  # 43.7 % of 16:9 pixels miss the open box
    # 16 sub-samples per pixel on a 64-row band through the middle (the oracle needs ~1 s for it)
    fp.aaEnabled, fp.aaSamples = 1, 4
    ctx.set_frame(fp)
    y0, y1 = 1048, 1112
    got = ctx.rt_draw(y0, y1)
    want = oracle.rt_draw(tris, fp, w, h, y0, y1)
    for k in ("pixelColours", "focalDistances", "closest"):
        assert np.array_equal(got[k][y0:y1].view(np.uint8), want[k][y0:y1].view(np.uint8)), k
    ctx.close()


def test_stats_count_rays(pkg, oracle):
    w, h = 96, 64
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 3
    ctx = pkg.Context(w, h)
    ctx.enable_stats(True)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    ctx.rt_draw()
    st = ctx.stats()
    want = oracle.rt_draw(tris, fp, w, h)
    assert st["primary_rays"] == want["primary_rays"] and st["shadow_rays"] == want["shadow_rays"]
    assert 0 < st["exact_tests"] < (st["primary_rays"] + st["shadow_rays"]) * len(tris)
    ctx.close()


def test_errors(pkg):
    ctx = pkg.Context(32, 32)
    with pytest.raises(pkg.B2RError):
        ctx.rt_draw()  # no scene yet
    ctx.set_triangles(pkg.cornell_box())
    with pytest.raises(pkg.B2RError):
        ctx.rt_draw()  # no frame params yet
    fp = pkg.default_frame_params(0, 32, 32)
    fp.numLights = 33
    with pytest.raises(pkg.B2RError):
        ctx.set_frame(fp)
    ctx.set_frame(pkg.default_frame_params(0, 32, 32))
    with pytest.raises(pkg.B2RError):
        ctx.rt_draw(5, 40)  # band outside the screen
    ctx.close()


@pytest.mark.parametrize("variant,k", [(0, 9), (2, 2), (3, 9), (0, 2), (3, 2), (1, 2)])
def test_large_scene_path(pkg, oracle, variant, k):
    """More than one 32-triangle chunk.  Scenes that do not fit in shared memory (k=9: 2,430 tessellated Cornell
    triangles) keep their per-frame constants in HBM; variant 2 forces that path for a small scene too; k=2 (120
    triangles) is the multi-chunk shared-memory layout; variant 3 = without the shadow-candidate cache.  AA and two
    lights."""
    w, h = 96, 64
    tris = pkg.tessellate(pkg.cornell_box(), k)
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 2
    fp.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14], [0.4, -0.2, -0.9, 0.3, 0.6, 0.9, 6]])
    ctx = pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RT_VARIANT, variant)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    want = oracle.rt_draw(tris, fp, w, h)
    for filt in (1, 0):
        ctx.set_option(pkg.capi.OPT_RT_FILTER, filt)
        check(ctx.rt_draw(), want)
    ctx.close()


def test_stl_sized_scene(pkg, oracle):
    """~9k triangles (the size of the reference's enemy1.stl): the brute-force semantics still hold."""
    w, h = 64, 40
    tris = pkg.tessellate(pkg.cornell_box(), 17)  # 8,670 triangles
    fp = pkg.default_frame_params(0, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    check(ctx.rt_draw(), oracle.rt_draw(tris, fp, w, h))
    ctx.close()


@pytest.mark.parametrize("aa,soft,nlights", [(0, 0, 1), (4, 0, 1), (3, 1, 1), (2, 0, 3)])
def test_direct_light_reuse_is_exact(pkg, oracle, aa, soft, nlights):
    """DirectLight is evaluated once per carried Intersection (a sub-sample whose hit does not replace the pixel's
    Intersection, raytracer.cpp:243, shades the same point as the one before).  Variant 5 evaluates it for every hit
    sub-sample like the reference: every array must be bit-identical either way, the ray counters must be the reference's
    in both, and only the 'evaluated' counter may differ."""
    rng = np.random.default_rng(100 + aa + 10 * soft + nlights)
    w, h = 200, 120
    tris = np.concatenate([pkg.cornell_box(), random_soup(rng, 12, spread=0.8, size=0.5)])
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled, fp.softShadowsSamples = int(aa > 0), max(aa, 1), soft, 4
    fp.set_camera([0.15, -0.1, -2.2], rot_y(0.2), h / 2)
    lights = np.array([[0, -0.5, -0.7, 1, 1, 1, 14], [0.6, -0.6, -0.2, 0.9, 0.7, 0.4, 6], [-0.5, 0.2, -0.9, 0.3, 0.5, 1.0, 9]], np.float32)[:nlights]
    fp.set_lights(lights)
    fp.set_random_positions(rng.uniform(-1, 1, (256, 3)).astype(np.float32))
    want = oracle.rt_draw(tris, fp, w, h)
    out = {}
    for variant in (0, 5):
        ctx = pkg.Context(w, h)
        ctx.set_option(pkg.capi.OPT_RT_VARIANT, variant)
        ctx.enable_stats(True)
        ctx.set_triangles(tris)
        ctx.set_frame(fp)
        got = ctx.rt_draw()
        st = ctx.stats()
        ctx.close()
        check(got, want)
        assert st["primary_rays"] == want["primary_rays"] and st["shadow_rays"] == want["shadow_rays"]
        out[variant] = st
    assert out[5]["shadow_rays_evaluated"] == out[5]["shadow_rays"]
    assert out[0]["shadow_rays_evaluated"] <= out[0]["shadow_rays"]
    if aa == 0:
        assert out[0]["shadow_rays_evaluated"] == out[0]["shadow_rays"]  # one sub-sample per pixel: nothing to reuse
    else:
        assert out[0]["shadow_rays_evaluated"] < out[0]["shadow_rays"]
