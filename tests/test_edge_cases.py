"""Edge cases of both paths against the CPU oracle (bit-exact unless stated)."""
import numpy as np
import pytest

from util import bits, random_soup, rot_y

pytestmark = pytest.mark.gpu


def rt_equal(got, want):
    assert np.array_equal(got["closest"]["triangleIndex"], want["closest"]["triangleIndex"])
    assert np.array_equal(got["closest"].view(np.uint8), want["closest"].view(np.uint8))
    assert np.array_equal(bits(got["pixelColours"]), bits(want["pixelColours"]))
    assert np.array_equal(bits(got["focalDistances"]), bits(want["focalDistances"]))


def ras_equal(got, want):
    assert np.array_equal(got["winner"], want["winner"])
    for k in ("depthBuffer", "pixelColours", "focalDistances"):
        assert np.array_equal(bits(got[k]), bits(want[k])), k


def test_rt_many_lights_soft_shadow_table_limit(pkg, oracle):
    """16 lights x 16 samples = the whole randomPositions[256] table (raytracer.cpp:84,286); 257 ray origins."""
    rng = np.random.default_rng(21)
    w, h = 48, 32
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    lights = np.concatenate([rng.uniform(-0.8, 0.8, (16, 3)), rng.uniform(0.2, 1, (16, 3)), rng.uniform(1, 6, (16, 1))], 1)
    fp.set_lights(lights.astype(np.float32))
    fp.softShadowsEnabled, fp.softShadowsSamples = 1, 16
    fp.set_random_positions(rng.uniform(-0.9, 0.9, (256, 3)).astype(np.float32))
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    rt_equal(ctx.rt_draw(), oracle.rt_draw(tris, fp, w, h))
    fp.numLights = 17  # 17 x 16 > 256: the reference would index past its table
    with pytest.raises(pkg.B2RError):
        ctx.set_frame(fp)
    ctx.close()


def test_rt_aa8_and_custom_ambient(pkg, oracle):
    w, h = 40, 30
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples = 1, 8
    fp.indirectLight[:] = [0.05, 0.3, 0.11]
    fp.dofFocalLength = 2.25
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    rt_equal(ctx.rt_draw(), oracle.rt_draw(tris, fp, w, h))
    ctx.close()


@pytest.mark.parametrize("seed", [5, 6])
def test_rt_camera_inside_the_scene_and_skewed_rotation(pkg, oracle, seed):
    """Camera inside the box, non-orthonormal cameraRot, light close to surfaces: many grazing and two-sided hits."""
    rng = np.random.default_rng(seed)
    w, h = 96, 72
    tris = np.concatenate([pkg.cornell_box(), random_soup(rng, 10, spread=0.6, size=0.5)])
    fp = pkg.default_frame_params(0, w, h)
    rot = rot_y(rng.uniform(-3, 3))
    rot[4] = 1.3
    rot[1] = 0.2  # shear
    fp.set_camera(rng.uniform(-0.5, 0.5, 3).astype(np.float32), rot, h / 3)
    fp.set_lights([[0.0, -0.95, 0.0, 1, 0.9, 0.8, 9], [0.9, 0.9, 0.9, 0.5, 0.5, 1, 4]])
    fp.aaEnabled, fp.aaSamples = 1, 2
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    want = oracle.rt_draw(tris, fp, w, h)
    for filt in (1, 0):
        ctx.set_option(pkg.capi.OPT_RT_FILTER, filt)
        rt_equal(ctx.rt_draw(), want)
    ctx.close()


def test_ras_close_camera_everything_clipped_somewhere(pkg, oracle):
    """Camera just outside the open face: spans leave the screen on every side, rows above and below the screen."""
    w, h = 200, 120
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera([0.2, 0.1, -1.6], rot_y(0.2, 1.01), float(h))
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    culled = ctx.ras_cull()
    assert np.array_equal(culled, oracle.ras_cull(tris, fp, w, h))
    ras_equal(ctx.ras_draw(), oracle.ras_draw(tris, culled, fp, w, h))
    ctx.close()


def test_ras_subpixel_and_degenerate_triangles(pkg, oracle):
    """Thousands of tiny triangles (most cover 0-2 pixels), zero-area ones, single-row ones, no lights."""
    rng = np.random.default_rng(9)
    w, h = 160, 100
    tris = random_soup(rng, 4000, spread=1.0, size=0.02)
    tris[:, [2, 5, 8]] += 1.5
    tris[10, 3:9] = np.tile(tris[10, 0:3], 2)      # zero area
    tris[11, [4, 7]] = tris[11, 1]                 # all vertices on one y: one row
    tris[12, 3:6] = tris[12, 0:3]                  # repeated vertex
    fp = pkg.default_frame_params(1, w, h)
    ctx = pkg.Context(w, h)
    ctx.enable_stats(True)
    ctx.set_triangles(tris)
    ctx.set_culled(np.zeros(len(tris), np.uint8))
    for nl in (1, 0):
        fp.numLights = nl
        ctx.set_frame(fp)
        got = ctx.ras_draw()
        want = oracle.ras_draw(tris, None, fp, w, h)
        ras_equal(got, want)
        st = ctx.stats()
        assert (st["ras_triangles"], st["ras_rows"], st["ras_depth_tests"]) == (want["triangles"], want["rows"], want["depth_tests"])
    ctx.close()


def test_ras_mixed_small_and_large_triangles_and_custom_reflectance(pkg, oracle):
    """Both pipeline branches in one frame (triangles of 1..200 rows), exact ties between duplicates of each kind."""
    rng = np.random.default_rng(10)
    w, h = 256, 200
    big = random_soup(rng, 12, spread=0.8, size=1.2)
    small = random_soup(rng, 600, spread=1.0, size=0.08)
    tris = np.concatenate([big, small, big[:4], small[:50]])
    tris[:, [2, 5, 8]] += 1.2
    fp = pkg.default_frame_params(1, w, h)
    fp.currentReflectance[:] = [0.9, 0.5, 1.0]
    fp.indirectLight[:] = [0.1, 0.2, 0.3]
    fp.set_lights([[0.3, -0.6, -0.4, 1, 1, 0.6, 11], [-0.7, 0.2, 0.1, 0.2, 0.9, 0.4, 5], [0, 0, 3, 1, 1, 1, 20]])
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_culled(np.zeros(len(tris), np.uint8))
    ctx.set_frame(fp)
    got = ctx.ras_draw()
    want = oracle.ras_draw(tris, None, fp, w, h)
    ras_equal(got, want)
    assert (got["winner"] < len(big) + len(small)).all()  # duplicates never win an exact tie
    # a second frame on the same context (key buffer is cleared by the previous shade pass)
    fp.set_camera([0.1, 0.0, -2.5], rot_y(-0.1, 1.01), float(h))
    ctx.set_frame(fp)
    ras_equal(ctx.ras_draw(), oracle.ras_draw(tris, None, fp, w, h))
    # and a band of a third one
    part = ctx.ras_draw(50, 120)
    assert np.array_equal(part["winner"][50:120], oracle.ras_draw(tris, None, fp, w, h)["winner"][50:120])
    ctx.close()


def test_dof_with_aa_and_soft_shadows(pkg, oracle):
    w, h = 96, 64
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(0, w, h)
    fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled, fp.dofEnabled, fp.dofKernelSize = 1, 2, 1, 1, 5
    fp.set_random_positions(pkg.jitter_table(3, [0, -0.5, -0.7]))
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    surf = ctx.rt_frame()
    want = oracle.rt_draw(tris, fp, w, h)
    assert np.array_equal(surf, oracle.resolve_surface(want["pixelColours"], want["focalDistances"], True, 5))
    ctx.close()


def test_stl_mesh_end_to_end(pkg, oracle, tmp_path):
    """an ASCII STL goes through the loader (reference semantics: scale -0.05, grey) and both paths,
    with the reference's CUSTOM_MODEL camera (rasteriser.cpp:109)."""
    rng = np.random.default_rng(12)
    soup = random_soup(rng, 700, spread=14.0, size=5.0)
    path = tmp_path / "mesh.stl"
    with open(path, "w") as f:
        f.write("solid mesh\n")
        for t in soup:
            f.write(" facet normal 0 0 0\n  outer loop\n")
            for k in range(3):
                f.write("  vertex %.6g %.6g %.6g\n" % tuple(t[3 * k:3 * k + 3]))
            f.write("  endloop\n endfacet\n\n")
        f.write("endsolid\n")
    tris = pkg.load_stl(str(path))
    assert tris.shape == (700, 15) and np.abs(tris[:, :9]).max() < 1.5
    w, h = 200, 150
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera([0, -0.5, -5.0], rot_y(0.0, 1.01), float(h))
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    culled = ctx.ras_cull()
    assert np.array_equal(culled, oracle.ras_cull(tris, fp, w, h))
    ras_equal(ctx.ras_draw(), oracle.ras_draw(tris, culled, fp, w, h))
    fr = pkg.default_frame_params(0, w, h)
    fr.set_camera([0, -0.5, -5.0], rot_y(0.0), float(h))
    ctx.set_frame(fr)
    rt_equal(ctx.rt_draw(), oracle.rt_draw(tris, fr, w, h))
    ctx.close()


@pytest.mark.gpu
def test_pinned_host_buffers(pkg):
    """b2r_pin_host_buffer: frames land in a page-locked caller buffer exactly as in a pageable one; pinning twice
    and unpinning something that was never pinned are not errors."""
    w, h = 320, 264
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(pkg.default_frame_params(0, w, h))
    want = ctx.rt_frame()
    buf = np.zeros((h, w), np.uint32)
    ctx.pin_host_buffer(buf)
    ctx.pin_host_buffer(buf)
    for _ in range(2):
        buf[:] = 0xDEADBEEF
        ctx.rt_frame(buf)
        assert np.array_equal(buf, want)
    ctx.unpin_host_buffer(buf)
    ctx.unpin_host_buffer(np.zeros(16, np.uint32))
    ctx.close()
