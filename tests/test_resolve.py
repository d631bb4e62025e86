"""CalculateDOF + PutPixelSDL + BMP payload parity (CUDA vs oracle)."""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("which", [0, 1])
@pytest.mark.parametrize("dof", [0, 1])
def test_resolve_surface_and_bmp(pkg, oracle, which, dof, tmp_path):
    w, h = 160, 120
    tris = pkg.cornell_box()
    fp = pkg.default_frame_params(which, w, h)
    fp.dofEnabled = dof
    ctx = pkg.Context(w, h)
    ctx.set_triangles(tris)
    ctx.set_frame(fp)
    if which == 0:
        out = ctx.rt_draw()
    else:
        ctx.ras_cull()
        out = ctx.ras_draw()
    surf = ctx.resolve_surface()
    want = oracle.resolve_surface(out["pixelColours"], out["focalDistances"], bool(dof), 8)
    assert np.array_equal(surf, want)
    assert not surf[0].any() and not surf[-1].any() and not surf[:, 0].any() and not surf[:, -1].any()
    bgr = ctx.resolve_bgr8()
    assert np.array_equal(bgr, oracle.surface_to_bgr8(want))
    frame = ctx.rt_frame() if which == 0 else ctx.ras_frame()
    assert np.array_equal(frame, want)
    path = str(tmp_path / "frame.bmp")
    pkg.write_bmp(path, bgr, w, h)
    raw = open(path, "rb").read()
    assert raw[:2] == b"BM" and len(raw) == 54 + len(bgr) and raw[54:] == bgr.tobytes()
    assert int.from_bytes(raw[18:22], "little") == w and int.from_bytes(raw[22:26], "little") == h
    ctx.close()


def test_bmp_row_padding(pkg, oracle):
    w, h = 33, 17  # 99 bytes per row -> padded to 100
    fp = pkg.default_frame_params(0, w, h)
    ctx = pkg.Context(w, h)
    ctx.set_triangles(pkg.cornell_box())
    ctx.set_frame(fp)
    out = ctx.rt_draw()
    bgr = ctx.resolve_bgr8()
    assert len(bgr) == 100 * h
    assert np.array_equal(bgr, oracle.surface_to_bgr8(oracle.resolve_surface(out["pixelColours"], None)))
    ctx.close()
