"""The reference's callee functions as CUDA entry points, each against vectors produced by the reference's own
function (tests/golden/sub_*.npz, written from oracle/_ref) -- bit-exact -- and against the oracle on fresh inputs."""
import os

import numpy as np
import pytest

from util import GOLDEN, bits, random_soup

pytestmark = pytest.mark.gpu


def test_closest_intersection_and_direct_light_golden(pkg):
    z = np.load(os.path.join(GOLDEN, "sub_rt.npz"))
    ctx = pkg.Context(96, 64)
    ctx.set_triangles(z["tris"])
    ctx.set_frame(pkg.default_frame_params(0, 96, 64))
    n = len(z["starts"])
    hit, clo, _ = ctx.closest_intersection(z["starts"], z["dirs"], is_light=(np.arange(n) % 2 == 1))
    assert np.array_equal(hit, z["hits"].astype(bool))
    assert clo.tobytes() == z["closest"].tobytes()  # position, distance, index of ClosestIntersection (raytracer.cpp:245-247)
    light = ctx.direct_light(clo[hit])
    assert np.array_equal(bits(light), bits(z["direct_light"][hit]))
    ctx.close()


def test_closest_intersection_running_state_and_focal(pkg, oracle):
    """The in/out Intersection (distance carried in, `>=` tie rule) and the focalDistances side effect (:243-249)."""
    rng = np.random.default_rng(31)
    tris = random_soup(rng, 50)
    fp = pkg.default_frame_params(0, 64, 64)
    fp.dofFocalLength = 0.75
    ctx = pkg.Context(64, 64)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    n = 300
    starts = rng.uniform(-0.2, 0.2, (n, 3)).astype(np.float32) + np.array([0, 0, -2.5], np.float32)
    dirs = rng.uniform(-0.6, 0.6, (n, 3)).astype(np.float32)
    dirs[:, 2] = 1.0
    prior = np.zeros(n, pkg.capi.INTERSECTION_DTYPE)
    prior["distance"] = rng.uniform(1.0, 4.0, n).astype(np.float32)  # something already closer for some rays
    prior["triangleIndex"] = 7
    hit, clo, foc = ctx.closest_intersection(starts, dirs, closest=prior)
    for k in range(n):
        h, c, f = oracle.rt_closest_intersection(tris, starts[k], dirs[k], closest=prior[k], dof_focal=0.75)
        assert h == hit[k] and c.tobytes() == clo[k].tobytes() and np.float32(f).tobytes() == foc[k].tobytes(), k
    ctx.close()


def test_rasteriser_stages_golden(pkg):
    z = np.load(os.path.join(GOLDEN, "sub_ras.npz"))
    w, h = 160, 120
    fp = pkg.default_frame_params(1, w, h)
    fp.set_camera(z["pos"], z["rot"], float(z["focal"]))
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    vp = ctx.vertex_shader(z["verts"])  # VertexShader, rasteriser.cpp:532
    assert vp.tobytes() == z["vertex_pixels"].tobytes()
    off = z["row_offsets"]
    for i in range(len(off) - 1):  # ComputePolygonRows, :674
        l, r = ctx.compute_polygon_rows(z["vertex_pixels"][3 * i:3 * i + 3])
        assert l.tobytes() == z["left"][off[i]:off[i + 1]].tobytes() and r.tobytes() == z["right"][off[i]:off[i + 1]].tobytes(), i
    ioff = z["interp_offsets"]
    for i in range(len(ioff) - 1):  # Interpolate, :615
        n = int(ioff[i + 1] - ioff[i])
        got = ctx.interpolate(z["vertex_pixels"][2 * i], z["vertex_pixels"][2 * i + 1], n)
        assert got.tobytes() == z["interp"][ioff[i]:ioff[i + 1]].tobytes(), i
    col, foc = ctx.pixel_shader(z["ps_in"], z["ps_col"], z["ps_nrm"])  # PixelShader, :549
    assert np.array_equal(bits(col), bits(z["ps_out"])) and np.array_equal(bits(foc), bits(z["ps_foc"]))
    with pytest.raises(pkg.B2RError):
        ctx.compute_polygon_rows(z["vertex_pixels"][:3], max_rows=1)
    ctx.close()
