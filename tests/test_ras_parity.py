"""Rasteriser parity: CUDA path (through the C ABI) vs the CPU oracle.

Criteria (north star): coverage and depth winner bit-exact, colour within 1e-4.  Depth values,
focal distances and colours are in fact bit-exact here and asserted as such.
"""
import numpy as np
import pytest

from util import bits, random_soup, rot_y

pytestmark = pytest.mark.gpu
COLOUR_TOL = 1e-4


def check(got, want):
    assert np.array_equal(got["winner"], want["winner"]), int((got["winner"] != want["winner"]).sum())
    assert np.array_equal(bits(got["depthBuffer"]), bits(want["depthBuffer"]))
    assert np.abs(got["pixelColours"] - want["pixelColours"]).max() <= COLOUR_TOL
    assert np.array_equal(bits(got["pixelColours"]), bits(want["pixelColours"]))
    assert np.array_equal(bits(got["focalDistances"]), bits(want["focalDistances"]))


@pytest.fixture(params=[0, 1, 2, 3], ids=["sortlast-slots", "sortlast-readback", "tiles-slots", "tiles-readback"])
def ras_variant(request):
    """Pipeline (B2R_OPT_RAS_VARIANT): sort-last (0, 1) or screen tiles (2, 3); large-triangle path with fixed-capacity
    slots and no readback (0, 2: small scenes) or with counters read back (1, 3)."""
    return request.param


def draw_both(pkg, oracle, tris, fp, w, h, cull=True, variant=0):
    ctx = pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, variant)
    ctx.enable_stats(True)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    if cull:
        culled = ctx.ras_cull()
        assert np.array_equal(culled, oracle.ras_cull(tris, fp, w, h))
    else:
        culled = np.zeros(len(tris), np.uint8)
        ctx.set_culled(culled)
    got = ctx.ras_draw()
    st = ctx.stats()
    ctx.close()
    want = oracle.ras_draw(tris, culled, fp, w, h)
    assert st["ras_triangles"] == want["triangles"] and st["ras_rows"] == want["rows"], (st, want["rows"])
    assert st["ras_depth_tests"] == want["depth_tests"], (st, want["depth_tests"])
    return got, want, culled


def test_config2_cornell_500(pkg, oracle, ras_variant):
    """BASELINE config 2: Cornell box 500x500, per-pixel illumination, 1/z depth buffer."""
    w = h = 500
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(1, w, h)
    got, want, culled = draw_both(pkg, oracle, tris, fp, w, h, variant=ras_variant)
    check(got, want)
    # known answers of the reference itself
    assert "".join(map(str, culled)) == "000000000000001111000011110011"
    assert int((got["winner"] >= 0).sum()) == 249498
    assert want["depth_tests"] == 291070


@pytest.mark.parametrize("w,h", [(96, 64), (64, 96), (160, 120), (33, 17)])
def test_small_screens(pkg, oracle, w, h, ras_variant):
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(1, w, h)
    got, want, _ = draw_both(pkg, oracle, tris, fp, w, h, variant=ras_variant)
    check(got, want)


@pytest.mark.parametrize("seed", [1, 2, 3, 4])
def test_random_scenes(pkg, oracle, seed, ras_variant):
    """Random soups in front of a rotated camera: off-screen spans, exact-zinv ties, several lights, no culling."""
    rng = np.random.default_rng(seed)
    w, h = 128, 96
    tris = random_soup(rng, 60, spread=1.2, size=0.9)
    tris[:, [2, 5, 8]] += 1.0  # keep every vertex in front of the camera
    if seed == 4:  # duplicate triangles: exact depth ties must go to the lower index
        tris[30:] = tris[:30]
        tris[30:, 12:15] = rng.uniform(0.1, 0.9, (30, 3)).astype(np.float32)
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera([0.1, -0.05, -3.0], rot_y(rng.uniform(-0.2, 0.2), 1.01), float(h))
    lights = np.concatenate([rng.uniform(-1, 1, (2, 3)), rng.uniform(0.2, 1, (2, 3)), rng.uniform(2, 20, (2, 1))], 1)
    fp.set_lights(lights.astype(np.float32))
    got, want, _ = draw_both(pkg, oracle, tris, fp, w, h, cull=(seed % 2 == 0), variant=ras_variant)
    check(got, want)
    if seed == 4:
        assert (got["winner"] < 30).all()


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_tessellated_cornell(pkg, oracle, variant):
    """BASELINE config 4 shape at reduced size: k=24 tessellation (17,280 triangles) at 640x360."""
    w, h = 640, 360
    tris = pkg.tessellate(pkg.cornell_box(), 24)
    fp = pkg.default_frame_params(1, w, h)
    got, want, _ = draw_both(pkg, oracle, tris, fp, w, h, variant=variant)
    check(got, want)


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_config4_full_size(pkg, oracle, variant):
    """BASELINE config 4 at full size: 183x183 tessellation = 1,004,670 triangles at 3840x2160."""
    w, h = 3840, 2160
    tris = pkg.tessellate(pkg.cornell_box(), 183)
    assert len(tris) == 1004670
    fp = pkg.default_frame_params(1, w, h)
    got, want, culled = draw_both(pkg, oracle, tris, fp, w, h, variant=variant)
    check(got, want)


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_row_bands_equal_full_frame(pkg, oracle, variant):
    w, h = 160, 120
    tris = pkg.tessellate(pkg.cornell_box(), 6)
    fp = pkg.default_frame_params(1, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, variant)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    ctx.ras_cull()
    full = ctx.ras_draw()
    merged = {k: np.zeros_like(v) for k, v in full.items()}
    merged["winner"][:] = -1
    for y0, y1 in ((0, 41), (41, 41), (41, 120)):
        part = ctx.ras_draw(y0, y1)
        for k in merged:
            merged[k][y0:y1] = part[k][y0:y1]
    for k in full:
        assert np.array_equal(full[k].view(np.uint8), merged[k].view(np.uint8)), k
    ctx.close()


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_empty_and_all_culled(pkg, oracle, variant):
    w, h = 64, 48
    ctx = pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, variant)
    ctx.set_triangles(np.zeros((0, 15), np.float32))
    ctx.set_frame(pkg.default_frame_params(1, w, h))
    got = ctx.ras_draw()
    assert (got["winner"] == -1).all() and not got["depthBuffer"].any() and not got["pixelColours"].any()
    tris = pkg.cornell_box()
    ctx.set_triangles(tris)
    ctx.set_culled(np.ones(len(tris), np.uint8))
    got = ctx.ras_draw()
    assert (got["winner"] == -1).all()
    ctx.close()


def test_vertex_behind_camera_is_refused(pkg, ras_variant):
    """A vertex on the camera plane projects to +-inf: the reference would walk ~2^31 rows; we return an error --
    from the host-buffer call itself, and for asynchronous draws of small scenes (no readback) from b2r_synchronize;
    a good frame afterwards is not affected."""
    import torch
    w, h = 64, 48
    good = pkg.cornell_box()
    tris = good[:1].copy()
    tris[0, 2] = -3.0  # v0.z == cameraPos.z
    ctx = pkg.Context(w, h)
    ctx.set_option(pkg.capi.OPT_RAS_VARIANT, ras_variant)
    ctx.set_triangles(tris)
    ctx.set_frame(pkg.default_frame_params(1, w, h))
    with pytest.raises(pkg.B2RError):
        ctx.ras_draw()
    col = torch.zeros((h, w, 3), dtype=torch.float32, device="cuda:0")
    with pytest.raises(pkg.B2RError):
        ctx.ras_draw_device_async(0, h, 0, col.data_ptr())
        ctx.synchronize()
    ctx.synchronize()  # reported once
    ctx.set_triangles(good)
    ctx.ras_cull()
    ctx.ras_draw()
    ctx.ras_draw_device_async(0, h, 0, col.data_ptr())
    ctx.synchronize()
    ctx.close()


@pytest.mark.parametrize("variant", [0, 2], ids=["sortlast", "tiles"])
def test_more_than_2_pow_20_triangles(pkg, oracle, variant):
    """1,083,000 sub-pixel triangles on a small screen: the tile pipeline's depth key holds 20 bits of the triangle index,
    so this scene is drawn in two epochs (indices below / from 2^20) with the keys frozen in between; tile lists run to
    tens of thousands of entries, i.e. many jobs per tile merged through the partial-result slots."""
    w, h = 320, 180
    tris = np.ascontiguousarray(pkg.tessellate(pkg.cornell_box(), 190)[::-1])  # reversed: the floor gets the highest indices
    assert len(tris) == 30 * 190 * 190 > (1 << 20)
    fp = pkg.default_frame_params(1, w, h)
    got, want, _ = draw_both(pkg, oracle, tris, fp, w, h, variant=variant)
    check(got, want)
    assert (got["winner"] >= (1 << 20)).any() and (got["winner"] < (1 << 20)).any()


@pytest.mark.parametrize("variant", [0, 1, 2, 3])
def test_wide_flat_triangles(pkg, oracle, variant):
    """Triangles of a few rows that span thousands of columns: in the tile pipeline they are wider than the 128-tile
    rectangle a small triangle may have and take the large-triangle path; long spans cross many tiles."""
    w, h = 6000, 48
    rng = np.random.default_rng(3)
    n = 24
    tris = np.zeros((n, 15), np.float32)
    for i in range(n):
        y = rng.uniform(-0.018, 0.018)
        z = rng.uniform(0.5, 3.0)
        v0 = [rng.uniform(-1.2, -0.6), y, z]
        v1 = [rng.uniform(0.6, 1.2), y + rng.uniform(-0.004, 0.004), z + rng.uniform(-0.3, 0.3)]
        v2 = [rng.uniform(-0.3, 0.3), y + rng.uniform(0.002, 0.01), z + rng.uniform(-0.3, 0.3)]
        tris[i, 0:9] = v0 + v1 + v2
        e1, e2 = tris[i, 3:6] - tris[i, 0:3], tris[i, 6:9] - tris[i, 0:3]
        nrm = np.cross(e2, e1)
        tris[i, 9:12] = nrm / np.linalg.norm(nrm)
        tris[i, 12:15] = rng.uniform(0.2, 0.9, 3)
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera([0.0, 0.0, -2.0], rot_y(0.0, 1.01), 2400.0)
    got, want, _ = draw_both(pkg, oracle, tris, fp, w, h, cull=False, variant=variant)
    check(got, want)
    assert (want["winner"] >= 0).sum() > 20000
