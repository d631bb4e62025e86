"""CPU oracle (oracle/liboracle.so) pinned against golden vectors produced by the reference itself.

tests/golden/*.npz were written by tests/golden/make_golden.py from oracle/_ref (the reference's own
raytracer.cpp / rasteriser.cpp compiled in the build container).  Bit-exact on every array.
"""
import json
import os

import numpy as np
import pytest

from util import GOLDEN, bits, fnv1a32, golden_files, ras_params_from_golden, rt_params_from_golden, size_from_name


@pytest.mark.parametrize("path", golden_files("rt_"), ids=os.path.basename)
def test_rt_frames(pkg, oracle, path):
    z = np.load(path)
    w, h = size_from_name(path)
    fp = rt_params_from_golden(pkg, z, w, h)
    o = oracle.rt_draw(z["tris"], fp, w, h)
    assert np.array_equal(o["closest"].view(np.uint8), z["closest"].view(np.uint8))
    assert np.array_equal(bits(o["pixelColours"]), bits(z["pixelColours"]))
    assert np.array_equal(bits(o["focalDistances"]), bits(z["focalDistances"]))
    assert np.array_equal(oracle.resolve_surface(o["pixelColours"], None), z["surface"])


@pytest.mark.parametrize("path", golden_files("ras_"), ids=os.path.basename)
def test_ras_frames(pkg, oracle, path):
    z = np.load(path)
    w, h = size_from_name(path)
    fp = ras_params_from_golden(pkg, z, w, h)
    assert np.array_equal(oracle.ras_cull(z["tris"], fp, w, h), z["culled"])
    o = oracle.ras_draw(z["tris"], z["culled"], fp, w, h)
    assert np.array_equal(o["winner"], z["winner"])
    for k in ("depthBuffer", "pixelColours", "focalDistances"):
        assert np.array_equal(bits(o[k]), bits(z[k])), k
    assert [o["depth_tests"], o["depth_passes"]] == z["counts"].tolist()
    assert np.array_equal(oracle.resolve_surface(o["pixelColours"], None), z["surface"])


def test_rt_substages(pkg, oracle):
    """ClosestIntersection / DirectLight with the reference signatures (raytracer.cpp:105-107)."""
    z = np.load(os.path.join(GOLDEN, "sub_rt.npz"))
    fp = pkg.default_frame_params(0, 96, 64)
    for i in range(len(z["starts"])):
        hit, c, _ = oracle.rt_closest_intersection(z["tris"], z["starts"][i], z["dirs"][i], is_light=(i % 2 == 1))
        assert hit == bool(z["hits"][i])
        assert c.tobytes() == z["closest"][i].tobytes()
        if hit:
            assert bits(oracle.rt_direct_light(z["tris"], fp, c)).tolist() == bits(z["direct_light"][i]).tolist()
    assert z["hits"].sum() > 100


def test_ras_substages(pkg, oracle):
    """VertexShader / Interpolate / ComputePolygonRows / PixelShader (rasteriser.cpp:532-735)."""
    z = np.load(os.path.join(GOLDEN, "sub_ras.npz"))
    w, h = 160, 120
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera(z["pos"], z["rot"], float(z["focal"]))
    vp = z["vertex_pixels"]
    for i in range(len(z["verts"])):
        assert oracle.ras_vertex_shader(fp, w, h, z["verts"][i]).tobytes() == vp[i].tobytes()
    off = z["row_offsets"]
    for i in range(len(off) - 1):
        l, r = oracle.ras_compute_polygon_rows(vp[3 * i:3 * i + 3])
        assert l.tobytes() == z["left"][off[i]:off[i + 1]].tobytes()
        assert r.tobytes() == z["right"][off[i]:off[i + 1]].tobytes()
    ioff = z["interp_offsets"]
    for i in range(len(ioff) - 1):
        n = int(ioff[i + 1] - ioff[i])
        assert oracle.ras_interpolate(vp[2 * i], vp[2 * i + 1], n).tobytes() == z["interp"][ioff[i]:ioff[i + 1]].tobytes()
    for i in range(len(z["ps_in"])):
        col, foc = oracle.ras_pixel_shader(fp, w, h, z["ps_in"][i], z["ps_col"][i], z["ps_nrm"][i])
        assert bits(col).tolist() == bits(z["ps_out"][i]).tolist()
        assert np.float32(foc).tobytes() == z["ps_foc"][i].tobytes()


def test_known_answers_500(pkg, oracle):
    """The reference's default frames: digests and the known answers listed in SURVEY.md section 7."""
    import hashlib
    k = json.load(open(os.path.join(GOLDEN, "kat_500x500.json")))
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    tris = pkg.cornell_box()
    assert "%08x" % fnv1a32(tris.tobytes()) == k["scene_fnv1a32"] == "b715a8a2"
    o = oracle.rt_draw(tris, pkg.default_frame_params(0, 500, 500), 500, 500)
    idx = o["closest"]["triangleIndex"]
    assert {str(int(a)): int(b) for a, b in zip(*np.unique(idx, return_counts=True))} == k["rt"]["hit_histogram"]
    for n in ("pixelColours", "focalDistances", "closest"):
        assert sha(o[n]) == k["rt"]["sha256"][n], n
    surf = oracle.resolve_surface(o["pixelColours"], None)
    assert sha(surf) == k["rt"]["sha256"]["surface"]
    assert int(((surf & 0xFFFFFF) == 0).sum()) == k["rt"]["black_surface_pixels"] == 1996
    fp = pkg.default_frame_params(1, 500, 500)
    culled = oracle.ras_cull(tris, fp, 500, 500)
    assert "".join(map(str, culled)) == k["ras"]["culled"] == "000000000000001111000011110011"
    o = oracle.ras_draw(tris, culled, fp, 500, 500)
    assert int((o["winner"] >= 0).sum()) == k["ras"]["covered"] == 249498
    assert o["depth_tests"] == k["ras"]["depth_tests"] == 291070 and o["depth_passes"] == k["ras"]["depth_passes"]
    for n in ("depthBuffer", "pixelColours", "focalDistances", "winner"):
        assert sha(o[n]) == k["ras"]["sha256"][n], n
    assert sha(oracle.resolve_surface(o["pixelColours"], None)) == k["ras"]["sha256"]["surface"]
