"""CPU oracle vs the reference itself (oracle/_ref/*.so), live, on seeded random inputs.

The .so files are built from /root/reference by oracle/build_ref.py in the build container and
travel to the GPU box prebuilt; where they are absent these tests skip (the golden fixtures in
test_oracle_golden.py still pin the oracle).
"""
import numpy as np
import pytest

from oracle import refbind
from util import bits, random_soup, rot_y

needs_ref = pytest.mark.skipif(not refbind.available("rt", 96, 64), reason="oracle/_ref not built")


@needs_ref
@pytest.mark.parametrize("w,h", [(96, 64), (64, 96)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_rt_random_scenes(pkg, oracle, w, h, seed):
    rng = np.random.default_rng(100 + seed)
    rt = refbind.RefRaytracer(w, h)
    tris = random_soup(rng, 20 + 10 * seed) if seed else rt.load_test_model()
    lights = np.concatenate([rng.uniform(-1, 1, (2, 3)), rng.uniform(0.2, 1, (2, 3)), rng.uniform(2, 20, (2, 1))], 1).astype(np.float32)
    table = rng.uniform(-1, 1, (256, 3)).astype(np.float32)
    pos = np.array([0.1 * seed, -0.05, -2.2], np.float32)
    rot = rot_y(0.2 * seed)
    rt.set_triangles(tris)
    rt.set_lights(lights, table)
    rt.set_camera(pos, rot, h / 2.0)
    fp = pkg.default_frame_params(0, w, h)
    fp.set_camera(pos, rot, h / 2.0).set_lights(lights).set_random_positions(table)
    fp.softShadowsSamples = 5
    for aa, soft in ((0, 0), (3, 0), (0, 1), (2, 1)):
        rt.set_flags(aa=aa > 0, aa_samples=max(aa, 1), soft=bool(soft), soft_samples=5)
        fp.aaEnabled, fp.aaSamples, fp.softShadowsEnabled = int(aa > 0), max(aa, 1), soft
        r = rt.draw()
        o = oracle.rt_draw(tris, fp, w, h)
        assert np.array_equal(r["closest"].view(np.uint8), o["closest"].view(np.uint8))
        assert np.array_equal(bits(r["pixelColours"]), bits(o["pixelColours"]))
        assert np.array_equal(bits(r["focalDistances"]), bits(o["focalDistances"]))
        assert np.array_equal(r["surface"], oracle.resolve_surface(o["pixelColours"], None))


@needs_ref
@pytest.mark.parametrize("w,h", [(96, 64), (160, 120)])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_ras_random_scenes(pkg, oracle, w, h, seed):
    rng = np.random.default_rng(200 + seed)
    ra = refbind.RefRasteriser(w, h)
    if seed:
        tris = random_soup(rng, 50, spread=1.2, size=0.9)
        tris[:, [2, 5, 8]] += 1.0
        ra.set_triangles(tris)
    else:
        tris = ra.load_test_model()
    lights = np.concatenate([rng.uniform(-1, 1, (2, 3)), rng.uniform(0.2, 1, (2, 3)), rng.uniform(2, 20, (2, 1))], 1).astype(np.float32)
    ra.set_lights(lights)
    ra.set_flags(backface=True, frustum=True)
    pos = np.array([0.2 * seed, 0.0, -3.0], np.float32)
    rot, culled, _ = ra.update_yaw(pos, 0.15 * seed, float(h))  # the reference's own Update(): rot + isCulled
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera(pos, rot, float(h)).set_lights(lights)
    assert np.array_equal(oracle.ras_cull(tris, fp, w, h), culled)
    r = ra.draw()
    o = oracle.ras_draw(tris, culled, fp, w, h)
    assert np.array_equal(r["winner"], o["winner"])
    for k in ("depthBuffer", "pixelColours", "focalDistances"):
        assert np.array_equal(bits(r[k]), bits(o[k])), k
    assert (r["depth_tests"], r["depth_passes"]) == (o["depth_tests"], o["depth_passes"])


@needs_ref
def test_dof_resolve_matches_reference_interior(pkg, oracle):
    """CalculateDOF's blur (raytracer.cpp:623-640): equal to the reference wherever its 8x8 window stays inside the
    pixel array (rows 4..H-4); nearer the top/bottom the reference reads outside its array (undefined)."""
    w, h = 160, 120
    rt = refbind.RefRaytracer(w, h)
    rt.load_test_model()
    rt.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14]])
    rt.set_camera_yaw([0, 0, -2], 0.0, h / 2.0)
    rt.set_flags(dof=True, dof_focal=1.3)
    r = rt.draw()
    mine = oracle.resolve_surface(r["pixelColours"], r["focalDistances"], True, 8)
    assert np.array_equal(mine[5:h - 5], r["surface"][5:h - 5])


@needs_ref
def test_stl_mesh_through_reference_loader(pkg, oracle):
    """the reference's LoadSTL path (enemy1.stl, 9,028 facets) rendered by reference and oracle."""
    import os
    if not os.path.isdir("/root/reference/rasteriser"):
        pytest.skip("reference sources not mounted")
    w, h = 160, 120
    ra = refbind.RefRasteriser(w, h)
    tris = ra.load_stl("/root/reference/rasteriser")
    assert len(tris) == 9028
    ra.set_lights([[0, -0.5, -0.7, 1, 1, 1, 14]])
    ra.set_flags()
    pos = np.array([0, -0.5, -5.0], np.float32)  # rasteriser.cpp:109
    rot, culled, _ = ra.update_yaw(pos, 0.0, float(h))
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera(pos, rot, float(h))
    assert np.array_equal(oracle.ras_cull(tris, fp, w, h), culled)
    r = ra.draw()
    o = oracle.ras_draw(tris, culled, fp, w, h)
    assert np.array_equal(r["winner"], o["winner"]) and np.array_equal(bits(r["depthBuffer"]), bits(o["depthBuffer"]))
    assert np.array_equal(bits(r["pixelColours"]), bits(o["pixelColours"]))
