"""CUDA path vs the golden vectors written by the reference itself (tests/golden, no oracle in between)."""
import os

import numpy as np
import pytest

from util import bits, golden_files, ras_params_from_golden, rt_params_from_golden, size_from_name

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("path", golden_files("rt_"), ids=os.path.basename)
def test_rt_frames(pkg, path):
    z = np.load(path)
    w, h = size_from_name(path)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(z["tris"])
    ctx.set_frame(rt_params_from_golden(pkg, z, w, h))
    got = ctx.rt_draw()
    assert np.array_equal(got["closest"]["triangleIndex"], z["closest"]["triangleIndex"])  # hit index: bit-exact
    assert np.abs(got["pixelColours"] - z["pixelColours"]).max() <= 1e-4                   # north-star tolerance
    assert np.array_equal(got["closest"].view(np.uint8), z["closest"].view(np.uint8))
    assert np.array_equal(bits(got["pixelColours"]), bits(z["pixelColours"]))
    assert np.array_equal(bits(got["focalDistances"]), bits(z["focalDistances"]))
    assert np.array_equal(ctx.resolve_surface(), z["surface"])
    ctx.close()


@pytest.mark.parametrize("path", golden_files("ras_"), ids=os.path.basename)
def test_ras_frames(pkg, path):
    z = np.load(path)
    w, h = size_from_name(path)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(z["tris"])
    ctx.set_frame(ras_params_from_golden(pkg, z, w, h))
    assert np.array_equal(ctx.ras_cull(), z["culled"])  # the reference's own Update() produced these flags
    got = ctx.ras_draw()
    assert np.array_equal(got["winner"], z["winner"])
    assert np.array_equal(bits(got["depthBuffer"]), bits(z["depthBuffer"]))
    assert np.abs(got["pixelColours"] - z["pixelColours"]).max() <= 1e-4
    assert np.array_equal(bits(got["pixelColours"]), bits(z["pixelColours"]))
    assert np.array_equal(bits(got["focalDistances"]), bits(z["focalDistances"]))
    assert np.array_equal(ctx.resolve_surface(), z["surface"])
    ctx.close()
