"""End-to-end drop-in proof: the reference programs themselves (their main(), Update(), LoadTestModel,
SDLauxiliary.h, compiled from /root/reference by oracle/build_ref.py --integration) with ONLY the body of Draw()
replaced by the calls of INTEGRATION.md.  Their own main loop runs one iteration (B2R_STUB_FRAMES=1; the rasteriser's Update()
clears the surface on every later iteration without redrawing, rasteriser.cpp:183-192,140-144), their own
SDL_SaveBMP call writes screenshot.bmp, and that file must equal the frame of the unmodified reference arithmetic
(the CPU oracle) byte for byte."""
import os
import subprocess

import numpy as np
import pytest

from util import ROOT

pytestmark = pytest.mark.gpu
REF = os.path.join(ROOT, "oracle", "_ref")


@pytest.mark.parametrize("prog,w,h", [("raytracer", 160, 120), ("rasteriser", 160, 120), ("raytracer", 500, 500),
                                      ("rasteriser", 500, 500)])
def test_reference_main_with_b2r_draw(pkg, oracle, prog, w, h, tmp_path):
    exe = os.path.join(REF, f"{prog}_b2r_{w}x{h}")
    if not os.path.exists(exe):
        pytest.skip("integration binary not built (needs /root/reference at build time)")
    env = dict(os.environ, B2R_STUB_FRAMES="1")
    out = subprocess.run([exe], cwd=str(tmp_path), env=env, capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1000:] + out.stderr[-2000:]
    assert "b2r:" not in out.stderr
    raw = open(tmp_path / "screenshot.bmp", "rb").read()
    tris = pkg.cornell_box()
    if prog == "raytracer":
        # the reference's start-up state at this screen size: focalLength stays 250 (raytracer.cpp:69)
        fp = pkg.default_frame_params(0, w, h)
        fp.focalLength = 250.0
        col = oracle.rt_draw(tris, fp, w, h)["pixelColours"]
    else:
        fp = pkg.default_frame_params(1, w, h)
        fp.focalLength = 500.0  # rasteriser.cpp:41
        col = oracle.ras_draw(tris, oracle.ras_cull(tris, fp, w, h), fp, w, h)["pixelColours"]
    want = oracle.surface_to_bgr8(oracle.resolve_surface(col, None))
    assert len(raw) == 54 + len(want)
    assert raw[54:] == want.tobytes()
    assert np.frombuffer(raw[54:], np.uint8).any()
